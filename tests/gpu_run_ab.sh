#!/bin/bash
# developer tool (GPU box): A/B on one box — x1 from the shared table and early start-up loads in the translated streaming kernel (cfg2)
T=${1:-r02ah}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; }
B="timeout 300 python bench.py --warmup 5 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg"
for rep in 1 2 3; do
for x in 0 1; do for e in 0 1; do
run x${x}e${e}_20_$rep FX8010_TR_X1TAB=$x FX8010_TR_EARLY=$e $B --steps 20
done; done; done
for x in 0 1; do for e in 0 1; do
run x${x}e${e}_200 FX8010_TR_X1TAB=$x FX8010_TR_EARLY=$e $B --steps 200 --warmup 20
done; done
