#!/bin/bash
# developer tool (GPU box): translated serial kernel with the cp.async input ring — parity, cfg5, and cfg4 (self recurrence) against the instruction-major kernel
T=${1:-r02x}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -5 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" > $O/${T}_$tag.json 2>$O/${T}_$tag.err; summ "$tag" $O/${T}_$tag.json; }
B="timeout 300 python bench.py --warmup 3 --no-cpu-baseline --no-sharded --no-e2e"
run cfg5 $B --config cfg5 --steps 3
run cfg5_ring4 FX8010_TR_RING=4 $B --config cfg5 --steps 3
run cfg4_im $B --config cfg4 --steps 10
run cfg4_tr FX8010_TR_RECUR=1 $B --config cfg4 --steps 10
run cfg4_tr_K1 FX8010_TR_RECUR=1 FX8010_TR_K=1 $B --config cfg4 --steps 10
run cfg4_tr_ring64 FX8010_TR_RECUR=1 FX8010_TR_RING=64 $B --config cfg4 --steps 10
run cfg4_tr_B64 FX8010_TR_RECUR=1 FX8010_TUNE_B=64 $B --config cfg4 --steps 10
run cfg4_tr_K2 FX8010_TR_RECUR=1 FX8010_TR_K=2 $B --config cfg4 --steps 10
run cfg4s_im $B --config cfg4 --instances 8192 --steps 10
run cfg4s_tr FX8010_TR_RECUR=1 $B --config cfg4 --instances 8192 --steps 10
run cfg4s_tr_K1 FX8010_TR_RECUR=1 FX8010_TR_K=1 $B --config cfg4 --instances 8192 --steps 10
run cfg4s_tr_B64 FX8010_TR_RECUR=1 FX8010_TUNE_B=64 $B --config cfg4 --instances 8192 --steps 10
