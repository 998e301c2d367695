#!/bin/bash
# developer tool (GPU box): time-cut delay lines on the translated streaming kernel — parity, cfg3 at its four ring sizes
T=${1:-r02ak}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -15 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; tail -2 $O/${T}_$tag.err; }
B="timeout 300 python bench.py --warmup 3 --steps 10 --no-cpu-baseline --no-sharded --no-e2e --config cfg3"
for s in 100 1000 8192 65536; do run cfg3_$s $B --itram $s; done
