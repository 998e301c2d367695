#!/bin/bash
# developer tool (GPU box): instruction-major kernel after the host-provided ring pointer bases — TRAM parity modes, cfg3 with the translator off
T=${1:-r02am}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; tail -2 $O/${T}_$tag.err; }
B="timeout 300 python bench.py --warmup 3 --steps 10 --no-cpu-baseline --no-sharded --no-e2e --config cfg3"
for s in 1000 8192; do run cfg3_${s}_im FX8010_BENCH_TRANSLATE=0 $B --itram $s; done
timeout 200 python tests/fuzz_campaign.py 100 > $O/${T}_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -2 $O/${T}_fuzz.log
