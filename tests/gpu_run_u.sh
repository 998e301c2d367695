#!/bin/bash
# developer tool (GPU box): the program translator — parity tests, the whole GPU suite, cfg5 translated vs interpreted, the default bench line
T=${1:-r02u}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -15 $O/${T}_pytest_translate.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${T}_smoke.log
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_cfg5_translated.json 2> $O/${T}_cfg5_translated.err; echo "cfg5 translated rc=$?"; summ cfg5t $O/${T}_cfg5_translated.json; tail -3 $O/${T}_cfg5_translated.err
FX8010_BENCH_TRANSLATE=0 timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu-baseline > $O/${T}_cfg5_interp.json 2> $O/${T}_cfg5_interp.err; echo "cfg5 interp rc=$?"; summ cfg5i $O/${T}_cfg5_interp.json; tail -3 $O/${T}_cfg5_interp.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; summ bench $O/${T}_bench.json; tail -3 $O/${T}_bench.err
