#!/bin/bash
# developer tool (GPU box): cfg2 variants after the pair fusion / 8 table replicas + one ncu capture of the general interpreter on cfg5
T=${1:-r02b}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
tests/pipe_peaks > $O/${T}_pipe_peaks.json 2> $O/${T}_pipe_peaks.err; echo "pipe_peaks rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${T}_pytest.log
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --steps 20 --warmup 5 $Q > $O/${T}_cfg2_20.json 2>&1; summ cfg2_20 $O/${T}_cfg2_20.json
python bench.py --steps 200 --warmup 20 $Q --no-parity > $O/${T}_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_cfg2_200.json
FX8010_NO_PAIRS=1 python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_nopairs.json 2>&1; summ cfg2_nopairs $O/${T}_cfg2_nopairs.json
for b in 32 128; do FX8010_TUNE_B=$b python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_B$b.json 2>&1; summ cfg2_B$b $O/${T}_cfg2_B$b.json; done
for m in 4 16; do FX8010_TUNE_M=$m python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_M$m.json 2>&1; summ cfg2_M$m $O/${T}_cfg2_M$m.json; done
FX8010_TUNE_B=128 FX8010_TUNE_M=16 python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_B128M16.json 2>&1; summ cfg2_B128M16 $O/${T}_cfg2_B128M16.json
FX8010_TUNE_K=2 python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_K2.json 2>&1; summ cfg2_K2 $O/${T}_cfg2_K2.json
python bench.py --config cfg1 --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg1.json 2>&1; summ cfg1 $O/${T}_cfg1.json
python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_probe_cfg5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fx_interp -s 2 -c 1 -o $O/${T}_ncu_cfg5 python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_ncu_cfg5.log 2>&1; echo "ncu rc=$?"
