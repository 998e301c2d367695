#!/bin/bash
# developer tool (GPU box): segment length of the translated streaming kernel against wave quantisation (cfg2, 20 blocks per launch)
T=${1:-r02ac}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; }
B="timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg"
for sl in 32 28 24 20 36 44 56 12 16 32; do run sl$sl FX8010_TUNE_SEGLEN=$sl $B; done
