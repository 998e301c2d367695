#!/bin/bash
T=${1:-r02k}; G=${2:-2}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
for i in 1 2 3 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus $G --steps 20 --warmup 5 --no-sharded > $O/${T}_bench_n${G}_$i.json 2> $O/${T}_bench_n${G}_$i.err; echo "bench N=$G run $i rc=$?"; summ bench_n$G $O/${T}_bench_n${G}_$i.json; grep -v "^$" $O/${T}_bench_n${G}_$i.err | grep -A30 "Fatal Python\|Segmentation" | head -50
done
