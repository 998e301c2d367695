#!/bin/bash
# developer tool (GPU box): what the driver runs at round end — the GPU suite, smoke(), bench.py both arms at the driver's flags — plus the translator's fuzz
T=${1:-r02final}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${T}_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${T}_ref.json 2> $O/${T}_ref.err; echo "reference rc=$?"; cut -c1-200 $O/${T}_ref.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"; summ bench $O/${T}_bench.json; tail -2 $O/${T}_bench.err
timeout 600 python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "bench default rc=$?"; summ default $O/${T}_bench_default.json
timeout 300 python tests/fuzz_campaign.py 150 translate > $O/${T}_fuzz_translate.log 2>&1; echo "fuzz translate rc=$?"; tail -2 $O/${T}_fuzz_translate.log
