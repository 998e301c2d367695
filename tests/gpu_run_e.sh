#!/bin/bash
# developer tool (GPU box): time-split delay lines (cfg3), INTERP loop copies (cfg4), ncu of the cfg2 fused launch (steady-state DRAM traffic)
T=${1:-r02e}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
for s in 100 1000 8192 65536; do python bench.py --config cfg3 --itram $s --steps 20 --warmup 5 $Q > $O/${T}_cfg3_$s.json 2>&1; summ cfg3_$s $O/${T}_cfg3_$s.json; done
for v in "2 8" "2 16" "4 8" "4 16" "1 16"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_M=$2 python bench.py --config cfg3 --itram 8192 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_K$1M$2.json 2>&1; summ cfg3_8192_K$1M$2 $O/${T}_cfg3_K$1M$2.json; done
for v in "2 16" "4 8"; do set -- $v; FX8010_TUNE_K=$1 FX8010_TUNE_M=$2 python bench.py --config cfg3 --itram 100 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg3_100_K$1M$2.json 2>&1; summ cfg3_100_K$1M$2 $O/${T}_cfg3_100_K$1M$2.json; done
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_cfg4.json 2>&1; summ cfg4 $O/${T}_cfg4.json
FX8010_TUNE_K=2 python bench.py --config cfg4 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg4_K2.json 2>&1; summ cfg4_K2 $O/${T}_cfg4_K2.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_cfg4_8192.json
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 3 -c 1 -o $O/${T}_ncu_cfg2_fused python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_ncu_cfg2.log 2>&1; echo "ncu rc=$?"
summ cfg2_20 $O/${T}_cfg2_20.json
