#!/bin/bash
# developer tool (GPU box, G GPUs): bench.py under torchrun, both arms, the way the driver launches them
T=${1:-r02m}; G=${2:-2}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus $G --steps 3 --warmup 1 > $O/${T}_ref_n$G.json 2> $O/${T}_ref_n$G.err; echo "reference N=$G rc=$?"; cut -c1-200 $O/${T}_ref_n$G.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 20 --warmup 5 > $O/${T}_bench_n$G.json 2> $O/${T}_bench_n$G.err; echo "bench N=$G rc=$?"; summ bench_n$G $O/${T}_bench_n$G.json; grep -v "^$" $O/${T}_bench_n$G.err | grep -A25 "Fatal Python\|Segmentation\|Error" | head -40
python tests/multi_bench.py $G > $O/${T}_multi_exec_n$G.json 2> $O/${T}_multi_exec_n$G.err; echo "multi_bench rc=$?"; cat $O/${T}_multi_exec_n$G.json
