#!/bin/bash
# developer tool (GPU box): delay lines after READ skipping + WRITE fusion; compute-sanitizer memcheck over the selected tests
T=${1:-r02f}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
for s in 100 1000 8192 65536; do python bench.py --config cfg3 --itram $s --steps 20 --warmup 5 $Q > $O/${T}_cfg3_$s.json 2>&1; summ cfg3_$s $O/${T}_cfg3_$s.json; done
for v in "2 8" "2 16" "4 8" "4 16" "4 4"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_M=$2 python bench.py --config cfg3 --itram 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_K$1M$2.json 2>&1; summ cfg3_8192_K$1M$2 $O/${T}_cfg3_K$1M$2.json; done
for sg in 4 8 16 32; do FX8010_TUNE_SEG=$sg python bench.py --config cfg3 --itram 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_seg$sg.json 2>&1; summ cfg3_8192_seg$sg $O/${T}_cfg3_seg$sg.json; done
FX8010_NO_TSPLIT=1 python bench.py --config cfg3 --itram 100 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_100_serial.json 2>&1; summ cfg3_100_serial $O/${T}_cfg3_100_serial.json
FX8010_NO_TSPLIT=1 python bench.py --config cfg3 --itram 1000 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_1000_serial.json 2>&1; summ cfg3_1000_serial $O/${T}_cfg3_1000_serial.json
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1; summ cfg2_20 $O/${T}_cfg2_20.json
bash tests/gpu_sanitize.sh memcheck $T
