#!/bin/bash
# developer tool (GPU box): last check of the tree — the GPU suite, smoke(), bench.py at the driver's flags
T=${1:-r02last}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/${T}_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; summ bench $O/${T}_bench.json; tail -2 $O/${T}_bench.err
