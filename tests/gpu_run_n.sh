#!/bin/bash
# developer tool (GPU box): two-sample table loop, broadcast input
T=${1:-r02n}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1; summ cfg2_20 $O/${T}_cfg2_20.json
python bench.py --steps 200 --warmup 20 $Q --no-parity > $O/${T}_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_cfg2_200.json
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_cfg1.json 2>&1; summ cfg1 $O/${T}_cfg1.json
python tests/multi_bench.py 1 > $O/${T}_multi_exec_n1.json 2> $O/${T}_multi_exec_n1.err; echo "multi_bench rc=$?"; cat $O/${T}_multi_exec_n1.json; tail -2 $O/${T}_multi_exec_n1.err
