#!/bin/bash
# developer tool (GPU box): full GPU test suite + the default bench line on the current build
T=${1:-r02q}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${T}_smoke.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; summ cfg2_20 $O/${T}_bench.json
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "bench default rc=$?"; summ cfg2_default $O/${T}_bench_default.json
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_ref.json 2>&1; echo "ref rc=$?"; tail -c 400 $O/${T}_ref.json
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_cfg4_8192.json
