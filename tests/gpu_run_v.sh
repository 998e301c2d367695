#!/bin/bash
# developer tool (GPU box): translated streaming kernel for stateless programs — parity, cfg2 bench translated vs instruction-major, segment sweep
T=${1:-r02v}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
timeout 600 python -m pytest tests/test_gpu_translate.py -x -q > $O/${T}_pytest_translate.log 2>&1; echo "pytest translate rc=$?"; tail -15 $O/${T}_pytest_translate.log
for steps in 20 200; do
timeout 300 python bench.py --steps $steps --warmup 5 --no-cpu-baseline --no-sharded > $O/${T}_cfg2_tr_$steps.json 2> $O/${T}_cfg2_tr_$steps.err; echo "cfg2 translated $steps rc=$?"; summ cfg2t$steps $O/${T}_cfg2_tr_$steps.json; tail -3 $O/${T}_cfg2_tr_$steps.err
done
FX8010_BENCH_TRANSLATE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded > $O/${T}_cfg2_im.json 2> $O/${T}_cfg2_im.err; echo "cfg2 instruction-major rc=$?"; summ cfg2im $O/${T}_cfg2_im.json
for sl in 8 16 32 64 128; do for B in 64 128 256; do
FX8010_TUNE_SEGLEN=$sl FX8010_TUNE_B=$B timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-e2e > $O/${T}_cfg2_sl${sl}_B$B.json 2>/dev/null; summ "seglen$sl B$B" $O/${T}_cfg2_sl${sl}_B$B.json
done; done
timeout 300 python bench.py --config cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-sharded > $O/${T}_cfg1_tr.json 2>/dev/null; summ cfg1t $O/${T}_cfg1_tr.json
