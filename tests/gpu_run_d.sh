#!/bin/bash
# developer tool (GPU box): multi-GPU executor tests + one ncu capture of the recurrence kernel (cfg4 at 8 192 instances)
T=${1:-r02d}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_cfg4_8192.json
for m in 8 16 64; do FX8010_TUNE_M=$m python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg4_8192_M$m.json 2>&1; summ cfg4_8192_M$m $O/${T}_cfg4_8192_M$m.json; done
FX8010_NO_STATELESS=1 python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg4_8192_short.json 2>&1; summ cfg4_8192_short $O/${T}_cfg4_8192_short.json
python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 3 $Q --no-parity > $O/${T}_cfg5.json 2>&1; summ cfg5 $O/${T}_cfg5.json
python tests/probe_cfg.py cfg4 8192 1024 3 > $O/${T}_probe_cfg4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 2 -c 1 -o $O/${T}_ncu_cfg4_8192 python tests/probe_cfg.py cfg4 8192 1024 3 > $O/${T}_ncu_cfg4.log 2>&1; echo "ncu rc=$?"
