#!/bin/bash
# developer tool (GPU box, 8 GPUs): bench.py at N = 8 (both arms) and N = 4, the multi-GPU executor with a broadcast input
T=${1:-r02t}; G=${2:-8}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $G --steps 20 --warmup 5 > $O/${T}_bench_n$G.json 2> $O/${T}_bench_n$G.err; echo "bench N=$G rc=$?"; summ bench_n$G $O/${T}_bench_n$G.json; grep -v "^$" $O/${T}_bench_n$G.err | grep -A25 "Fatal Python\|Segmentation" | head -40
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 20 --warmup 5 > $O/${T}_bench_n4.json 2> $O/${T}_bench_n4.err; echo "bench N=4 rc=$?"; summ bench_n4 $O/${T}_bench_n4.json
python tests/multi_bench.py $G > $O/${T}_multi_exec_n$G.json 2> $O/${T}_multi_exec_n$G.err; echo "multi_bench rc=$?"; cat $O/${T}_multi_exec_n$G.json
