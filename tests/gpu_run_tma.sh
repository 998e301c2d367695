#!/bin/bash
# developer tool (GPU box): bulk tensor copies (TMA) for the input stage of the delay-free instruction-major kernel, opt-in (FX8010_USE_TMA=1)
T=${1:-r02p}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
FX8010_USE_TMA=1 timeout 300 python -m pytest tests -m gpu -x -q -k "cfg4 or cfg2 or cfg1 or carried or stateless_kernel_modes or pairs or fuzz_instruction_major or null_input or tiny_instance" > $O/${T}_pytest_tma.log 2>&1; echo "pytest(TMA) rc=$?"; tail -4 $O/${T}_pytest_tma.log
Q="--no-cpu-baseline --no-sharded --no-e2e --no-parity"
for t in 0 1; do
FX8010_USE_TMA=$t timeout 200 python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_tma$t.json 2>&1; summ cfg4_tma$t $O/${T}_cfg4_tma$t.json
FX8010_USE_TMA=$t timeout 200 python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4s_tma$t.json 2>&1; summ cfg4_8192_tma$t $O/${T}_cfg4s_tma$t.json
FX8010_USE_TMA=$t FX8010_NO_FUSE=1 timeout 200 python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_nofuse_tma$t.json 2>&1; summ cfg2_nofuse_tma$t $O/${T}_cfg2_nofuse_tma$t.json
done
FX8010_USE_TMA=1 timeout 200 python bench.py --config cfg4 --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e > $O/${T}_cfg4_tma_parity.json 2>&1; summ cfg4_tma_parity $O/${T}_cfg4_tma_parity.json
