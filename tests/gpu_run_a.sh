#!/bin/bash
# developer tool (GPU box): round-2 measurement pass — tests, pipe peaks, bench lines for every config, launch-plan sweeps.
# usage: bash tests/gpu_run_a.sh <tag>      (writes gpurun_out/<tag>_*)
T=${1:-r02a}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_smi.txt 2>&1
tests/pipe_peaks > $O/${T}_pipe_peaks.json 2> $O/${T}_pipe_peaks.err; echo "pipe_peaks rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/${T}_pytest.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_cfg2_20.json 2> $O/${T}_bench_cfg2_20.err; echo "bench20 rc=$?"; summ cfg2_20 $O/${T}_bench_cfg2_20.json
Q="--no-cpu-baseline --no-sharded"
python bench.py --steps 200 --warmup 20 $Q > $O/${T}_bench_cfg2_200.json 2>&1; summ cfg2_200 $O/${T}_bench_cfg2_200.json
Q="--no-cpu-baseline --no-sharded --no-e2e"
for seg in 4 8 16 32; do FX8010_TUNE_SEG=$seg python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_seg$seg.json 2>&1; summ cfg2_seg$seg $O/${T}_cfg2_seg$seg.json; done
FX8010_NO_FUSE=1 python bench.py --steps 20 --warmup 5 $Q --no-parity > $O/${T}_cfg2_nofuse.json 2>&1; summ cfg2_nofuse $O/${T}_cfg2_nofuse.json
python bench.py --config cfg1 --steps 20 --warmup 5 $Q > $O/${T}_bench_cfg1.json 2>&1; summ cfg1 $O/${T}_bench_cfg1.json
for s in 100 1000 8192 65536; do python bench.py --config cfg3 --itram $s --steps 20 --warmup 5 $Q > $O/${T}_bench_cfg3_$s.json 2>&1; summ cfg3_$s $O/${T}_bench_cfg3_$s.json; done
python bench.py --config cfg4 --steps 20 --warmup 5 $Q > $O/${T}_bench_cfg4.json 2>&1; summ cfg4 $O/${T}_bench_cfg4.json
python bench.py --config cfg4 --instances 8192 --steps 20 --warmup 5 $Q > $O/${T}_bench_cfg4_8192.json 2>&1; summ cfg4_8192 $O/${T}_bench_cfg4_8192.json
python bench.py --config cfg5 --steps 3 --warmup 3 --repeats 3 $Q > $O/${T}_bench_cfg5.json 2>&1; summ cfg5 $O/${T}_bench_cfg5.json
for sub in 256 1024; do FX8010_TUNE_SUB=$sub python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-parity > $O/${T}_cfg2_sub$sub.json 2>&1; summ cfg2_sub$sub $O/${T}_cfg2_sub$sub.json; done
