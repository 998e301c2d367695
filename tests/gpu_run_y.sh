#!/bin/bash
# developer tool (GPU box): if-converted SKIPs / branch-free helpers in the translated kernels — parity, cfg5 and cfg2 timing
T=${1:-r02aa}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -5 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; }
B="timeout 300 python bench.py --warmup 3 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg"
run cfg5 $B --config cfg5 --steps 3
run cfg5_ifconv FX8010_TR_IFCONV=1 $B --config cfg5 --steps 3
run cfg5_B64 FX8010_TUNE_B=64 $B --config cfg5 --steps 3
run cfg5_262144 $B --config cfg5 --instances 262144 --steps 2 --repeats 2
run cfg2 $B --steps 20 --warmup 5
run cfg2_200 $B --steps 200 --warmup 20
run cfg1 $B --config cfg1 --steps 20 --warmup 5
