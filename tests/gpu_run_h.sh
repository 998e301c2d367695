#!/bin/bash
# developer tool (GPU box): segment-count sweep for the fused cfg2 launch (wave quantisation)
T=${1:-r02h}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
Q="--no-cpu-baseline --no-sharded --no-e2e --no-parity"
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_auto.json 2>&1; summ cfg2_auto $O/${T}_cfg2_auto.json
for sg in 3 6 12 16 19 32; do FX8010_TUNE_SEG=$sg python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_seg$sg.json 2>&1; summ cfg2_seg$sg $O/${T}_cfg2_seg$sg.json; done
python bench.py --steps 200 --warmup 20 $Q > $O/${T}_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_cfg2_200.json
for sg in 2 4 6 8; do FX8010_TUNE_SEG=$sg python bench.py --steps 200 --warmup 20 $Q > $O/${T}_cfg2_200_seg$sg.json 2>&1; summ cfg2_200_seg$sg $O/${T}_cfg2_200_seg$sg.json; done
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_cfg1.json 2>&1; summ cfg1 $O/${T}_cfg1.json
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "facade_multi or multi_executor or cfg2_timed" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
