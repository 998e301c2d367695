#!/bin/bash
# developer tool (GPU box): ONE compute-sanitizer tool per call (B200_PROFILING.md) over a small set of parity tests that
# covers the three kernel families, the sample split (P > 1), pair fusion, fused multi-block launches and PDL overlap.
# usage: bash tests/gpu_sanitize.sh memcheck|racecheck <tag>
TOOL=${1:-memcheck}
T=${2:-r02}
O=gpurun_out
mkdir -p $O
SEL='test_snippets or test_cfg1_testcode_shipped or (test_tram_instruction_major and (K1P4M16 or P2) and (s100 or wr_off or s3)) or (test_carried_recurrences and auto) or (test_producer_consumer_pairs and auto) or (test_random_programs and 0) or (test_multi_executor_shards_on_one_device and cfg3) or test_end_skipped_wraps_and_cap or (test_stateless_kernel_modes and M8)'
python -m pytest tests -m gpu -x -q -k "$SEL" > $O/${T}_sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 $O/${T}_sanitize_plain.log
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --log-file $O/${T}_sanitize_${TOOL}.log python -m pytest tests -m gpu -x -q -k "$SEL" > $O/${T}_sanitize_${TOOL}_pytest.log 2>&1; echo "$TOOL rc=$?"
tail -3 $O/${T}_sanitize_${TOOL}_pytest.log; tail -5 $O/${T}_sanitize_${TOOL}.log
