#!/bin/bash
# developer tool (GPU box): the translator in the differential fuzz campaign + the translate test file
T=${1:-r02aj}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -5 $O/${T}_pytest_translate.log
timeout 400 python tests/fuzz_campaign.py 300 translate > $O/${T}_fuzz_translate.log 2>&1; echo "fuzz translate rc=$?"; tail -5 $O/${T}_fuzz_translate.log
timeout 300 python tests/fuzz_campaign.py 120 > $O/${T}_fuzz_all.log 2>&1; echo "fuzz all rc=$?"; tail -3 $O/${T}_fuzz_all.log
