#!/bin/bash
# developer tool (GPU box): the evidence pass behind profiles/ — bench lines of every config, the ncu launch list of the
# bench command and one `ncu --set full` capture per kernel family (each only after the same command exited 0 without ncu).
T=${1:-r02}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${T}_ev_smi.txt 2>&1
tests/pipe_peaks > $O/${T}_ev_pipe_peaks.json 2>/dev/null
python bench.py --steps 20 --warmup 5 > $O/${T}_ev_bench_cfg2.json 2> $O/${T}_ev_bench_cfg2.err; echo "bench rc=$?"; summ cfg2_20 $O/${T}_ev_bench_cfg2.json
Q="--no-cpu-baseline --no-sharded"
python bench.py --steps 200 --warmup 20 $Q > $O/${T}_ev_bench_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_ev_bench_cfg2_200.json
Q="--no-cpu-baseline --no-sharded --no-e2e"
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_ev_bench_cfg1.json 2>&1; summ cfg1 $O/${T}_ev_bench_cfg1.json
for s in 100 1000 8192 65536; do python bench.py --config cfg3 --itram $s --steps 20 --warmup 5 $Q > $O/${T}_ev_bench_cfg3_$s.json 2>&1; summ cfg3_$s $O/${T}_ev_bench_cfg3_$s.json; done
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_ev_bench_cfg4.json 2>&1; summ cfg4 $O/${T}_ev_bench_cfg4.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_ev_bench_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_ev_bench_cfg4_8192.json
python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 3 $Q > $O/${T}_ev_bench_cfg5.json 2>&1; summ cfg5 $O/${T}_ev_bench_cfg5.json
python bench.py --config cfg5 --instances 262144 --steps 2 --warmup 3 --repeats 2 $Q --no-parity > $O/${T}_ev_bench_cfg5_262144.json 2>&1; summ cfg5_262144 $O/${T}_ev_bench_cfg5_262144.json
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e --no-parity"
$B > $O/${T}_ev_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ev_launches_cfg2.csv $B > $O/${T}_ev_ncu_launches.log 2>&1; echo "launch list rc=$?"
$B > $O/${T}_ev_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 3 -c 1 -o $O/${T}_ev_ncu_cfg2 $B > $O/${T}_ev_ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
python tests/probe_cfg.py cfg4 65536 1024 3 > $O/${T}_ev_probe4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 2 -c 1 -o $O/${T}_ev_ncu_cfg4 python tests/probe_cfg.py cfg4 65536 1024 3 > $O/${T}_ev_ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
python tests/probe_cfg.py cfg3 16384 1024 3 > $O/${T}_ev_probe3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_stateless -s 2 -c 1 -o $O/${T}_ev_ncu_cfg3 python tests/probe_cfg.py cfg3 16384 1024 3 > $O/${T}_ev_ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_ev_probe5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fx_interp -s 2 -c 1 -o $O/${T}_ev_ncu_cfg5 python tests/probe_cfg.py cfg5 32768 128 3 > $O/${T}_ev_ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
