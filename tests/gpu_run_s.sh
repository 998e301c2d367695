#!/bin/bash
# developer tool (GPU box): serial translated kernel, G periods per ring transaction — parity, cfg3, cfg4, cfg5
T=${1:-r02af}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -15 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; tail -2 $O/${T}_$tag.err; }
B="timeout 300 python bench.py --warmup 3 --steps 10 --no-cpu-baseline --no-sharded --no-e2e --no-interpreter-leg"
for s in 100 1000 8192; do run cfg3_${s}_tr FX8010_TR_RECUR=2 $B --config cfg3 --itram $s; done
for g in 1 2 4; do run cfg3_1000_G$g FX8010_TR_RECUR=2 FX8010_TR_G=$g $B --config cfg3 --itram 1000; done
run cfg3_1000_nopin FX8010_TR_RECUR=2 FX8010_TR_PIN=0 $B --config cfg3 --itram 1000
run cfg3_1000_ring16 FX8010_TR_RECUR=2 FX8010_TR_RING=16 $B --config cfg3 --itram 1000
run cfg4_tr_K1 FX8010_TR_RECUR=1 FX8010_TR_K=1 $B --config cfg4
run cfg4_tr_K2 FX8010_TR_RECUR=1 FX8010_TR_K=2 $B --config cfg4
run cfg4s_tr_K1 FX8010_TR_RECUR=1 FX8010_TR_K=1 $B --config cfg4 --instances 8192
run cfg4s_tr_K1_G1 FX8010_TR_RECUR=1 FX8010_TR_K=1 FX8010_TR_G=1 $B --config cfg4 --instances 8192
run cfg5 $B --config cfg5 --steps 3
