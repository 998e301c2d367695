#!/bin/bash
# developer tool (GPU box): geometry sweep of the one-pole recurrence (cfg4) at 65 536 and 8 192 instances
T=${1:-r02o}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
Q="--no-cpu-baseline --no-sharded --no-e2e --no-parity"
for v in "4 32" "2 32" "2 64" "1 32" "1 64" "1 128" "2 128"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_B=$2 python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_K$1B$2.json 2>&1; summ cfg4_K$1B$2 $O/${T}_cfg4_K$1B$2.json; done
for v in "1 32 64" "1 32 32" "2 32 64" "1 64 64"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_B=$2 FX8010_TUNE_M=$3 python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4s_K$1B$2M$3.json 2>&1; summ cfg4_8192_K$1B$2M$3 $O/${T}_cfg4s_K$1B$2M$3.json; done
for v in "1 32" "2 32" "1 64"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_B=$2 FX8010_TUNE_M=32 python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_K$1B$2M32.json 2>&1; summ cfg4_K$1B$2M32 $O/${T}_cfg4_K$1B$2M32.json; done
