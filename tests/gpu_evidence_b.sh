#!/bin/bash
# developer tool (GPU box): evidence pass after the program translator — whole GPU suite, bench lines, launch list, and one
# `ncu --set full` capture each of the translated streaming kernel (cfg2) and the translated serial kernel (cfg5),
# each only after the same command exited 0 without ncu.
T=${1:-r02z}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 900 python -m pytest tests -x -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/${T}_smoke.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_cfg2.json 2> $O/${T}_bench_cfg2.err; echo "bench rc=$?"; summ cfg2_20 $O/${T}_bench_cfg2.json; tail -2 $O/${T}_bench_cfg2.err
python bench.py > $O/${T}_bench_default.json 2> $O/${T}_bench_default.err; echo "bench default rc=$?"; summ default $O/${T}_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err; echo "reference rc=$?"; cut -c1-300 $O/${T}_bench_reference.json
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_bench_cfg1.json 2>&1; summ cfg1 $O/${T}_bench_cfg1.json
python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 3 $Q > $O/${T}_bench_cfg5.json 2>&1; summ cfg5 $O/${T}_bench_cfg5.json
python bench.py --config cfg5 --instances 262144 --steps 2 --warmup 3 --repeats 2 $Q > $O/${T}_bench_cfg5_262144.json 2>&1; summ cfg5_262144 $O/${T}_bench_cfg5_262144.json
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e --no-parity"
$B > $O/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches_cfg2.csv $B > $O/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
$B > $O/${T}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_translated_sl -s 3 -c 1 -o $O/${T}_ncu_cfg2 $B > $O/${T}_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
export FX8010_TRANSLATE=2
python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_probe5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_translated -s 2 -c 1 -o $O/${T}_ncu_cfg5 python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"; cat $O/${T}_probe5.log
ls -la $O/${T}_ncu_*.ncu-rep
