#!/bin/bash
# developer tool (GPU box): differential fuzz campaign + end-to-end check after the larger host sub-blocks
T=${1:-r02r}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python tests/fuzz_campaign.py 420 > $O/${T}_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -5 $O/${T}_fuzz.log


