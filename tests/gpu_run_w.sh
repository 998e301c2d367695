#!/bin/bash
# developer tool (GPU box): streaming-kernel variants (group size, software pipeline, shared-memory tables, segment length) on cfg2
T=${1:-r02w}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -5 $O/${T}_pytest_translate.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e > $O/${T}_$tag.json 2>/dev/null; summ "$tag" $O/${T}_$tag.json; }
run base FX8010_TR_U=4
run pf0 FX8010_TR_PF=0
run smem0 FX8010_TR_SMEM=0
run pf0smem0 FX8010_TR_PF=0 FX8010_TR_SMEM=0
for U in 1 2 4 8; do for sl in 8 16 32 64; do
run U${U}_sl$sl FX8010_TR_U=$U FX8010_TUNE_SEGLEN=$sl
done; done
run U2_sl32_B64 FX8010_TR_U=2 FX8010_TUNE_SEGLEN=32 FX8010_TUNE_B=64
run U2_sl32_B256 FX8010_TR_U=2 FX8010_TUNE_SEGLEN=32 FX8010_TUNE_B=256
run U4_sl32_B64 FX8010_TR_U=4 FX8010_TUNE_SEGLEN=32 FX8010_TUNE_B=64
run U4_sl32_B256 FX8010_TR_U=4 FX8010_TUNE_SEGLEN=32 FX8010_TUNE_B=256
