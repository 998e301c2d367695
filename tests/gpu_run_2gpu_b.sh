#!/bin/bash
T=${1:-r02j}; G=${2:-2}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 5 > $O/${T}_bench_n$G.json 2> $O/${T}_bench_n$G.err; echo "bench N=$G rc=$?"; summ bench_n$G $O/${T}_bench_n$G.json; grep -v "^$" $O/${T}_bench_n$G.err | grep -B2 -A25 "Fatal Python\|Segmentation" | head -60
