#!/bin/bash
# developer tool (GPU box): TRAM read streams in the translated serial kernel — parity, cfg3 through it (FX8010_TR_RECUR=2) against the instruction-major kernel
T=${1:-r02ad}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -15 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; tail -2 $O/${T}_$tag.err; }
B="timeout 300 python bench.py --warmup 3 --steps 10 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg --config cfg3"
for s in 100 1000 8192 65536; do
run cfg3_${s}_im $B --itram $s
run cfg3_${s}_tr FX8010_TR_RECUR=2 $B --itram $s
done
run cfg3_1000_tr_ring16 FX8010_TR_RECUR=2 FX8010_TR_RING=16 $B --itram 1000
run cfg3_1000_tr_B64 FX8010_TR_RECUR=2 FX8010_TUNE_B=64 $B --itram 1000
run cfg3_1000_tr_nostream FX8010_TR_RECUR=2 FX8010_TR_STREAMS=0 $B --itram 1000
