#!/bin/bash
# developer tool (GPU box): wide pre-decoded encoding; ncu of the time-split delay line (cfg3, itramsize 8192)
T=${1:-r02g}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -18 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1; summ cfg2_20 $O/${T}_cfg2_20.json
python bench.py --steps 200 --warmup 20 $Q --no-parity > $O/${T}_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_cfg2_200.json
FX8010_TUNE_M=16 python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_M16.json 2>&1; summ cfg2_M16 $O/${T}_cfg2_M16.json
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_cfg1.json 2>&1; summ cfg1 $O/${T}_cfg1.json
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_cfg4_8192.json
for s in 100 1000 8192; do python bench.py --config cfg3 --itram $s --steps 20 --warmup 5 $Q > $O/${T}_cfg3_$s.json 2>&1; summ cfg3_$s $O/${T}_cfg3_$s.json; done
for v in "2 8" "2 16" "4 8" "1 16" "1 32" "2 32"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_M=$2 python bench.py --config cfg3 --itram 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_K$1M$2.json 2>&1; summ cfg3_8192_K$1M$2 $O/${T}_cfg3_K$1M$2.json; done
FX8010_NO_TSPLIT=1 python bench.py --config cfg3 --itram 100 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_100_serial.json 2>&1; summ cfg3_100_serial $O/${T}_cfg3_100_serial.json
FX8010_TUNE_K=2 FX8010_TUNE_M=16 python tests/probe_cfg.py cfg3 16384 1024 3 > $O/${T}_probe_cfg3.log 2>&1 && \
FX8010_TUNE_K=2 FX8010_TUNE_M=16 ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 2 -c 1 -o $O/${T}_ncu_cfg3 python tests/probe_cfg.py cfg3 16384 1024 3 > $O/${T}_ncu_cfg3.log 2>&1; echo "ncu rc=$?"
