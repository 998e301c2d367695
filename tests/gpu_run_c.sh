#!/bin/bash
# developer tool (GPU box): after INTERP integer widening + predicate/accumulator specialisation of the general interpreter
T=${1:-r02c}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
for k in 1 2; do FX8010_TUNE_K=$k python bench.py --config cfg4 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg4_K$k.json 2>&1; summ cfg4_K$k $O/${T}_cfg4_K$k.json; done
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_cfg4_8192.json
python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 3 $Q > $O/${T}_cfg5.json 2>&1; summ cfg5 $O/${T}_cfg5.json
for k in 1 4; do FX8010_TUNE_K=$k python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 2 $Q --no-parity > $O/${T}_cfg5_K$k.json 2>&1; summ cfg5_K$k $O/${T}_cfg5_K$k.json; done
FX8010_TUNE_K=1 FX8010_TUNE_B=32 python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 2 $Q --no-parity > $O/${T}_cfg5_K1B32.json 2>&1; summ cfg5_K1B32 $O/${T}_cfg5_K1B32.json
FX8010_TUNE_K=2 FX8010_TUNE_B=32 python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 2 $Q --no-parity > $O/${T}_cfg5_K2B32.json 2>&1; summ cfg5_K2B32 $O/${T}_cfg5_K2B32.json
python bench.py --config cfg3 --steps 20 --warmup 5 $Q > $O/${T}_cfg3.json 2>&1; summ cfg3 $O/${T}_cfg3.json
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1; summ cfg2_20 $O/${T}_cfg2_20.json
