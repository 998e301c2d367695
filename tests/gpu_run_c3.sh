#!/bin/bash
# developer tool (GPU box): load-group size / software pipeline of the delay-line streaming kernel (cfg3, ring of 8 192 and 100)
T=${1:-r02c3}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; }
B="timeout 200 python bench.py --warmup 3 --steps 10 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg --config cfg3"
run base $B --itram 8192
run U2 FX8010_TR_U=2 $B --itram 8192
run PF0 FX8010_TR_PF=0 $B --itram 8192
run U2PF0 FX8010_TR_U=2 FX8010_TR_PF=0 $B --itram 8192
run U8PF0 FX8010_TR_U=8 FX8010_TR_PF=0 $B --itram 8192
run B256 FX8010_TUNE_B=256 $B --itram 8192
run B64 FX8010_TUNE_B=64 $B --itram 8192
run sl16 FX8010_TUNE_SEGLEN=16 $B --itram 8192
run sl32 FX8010_TUNE_SEGLEN=32 $B --itram 8192
run s100_base $B --itram 100
run s100_U2 FX8010_TR_U=2 $B --itram 100
