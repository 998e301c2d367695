#!/bin/bash
# developer tool (GPU box, 2+ GPUs): the multi-GPU executor on two devices and bench.py under torchrun
T=${1:-r02i}
G=${2:-2}
O=gpurun_out
mkdir -p $O
. tests/gpu_summ.sh
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "multi_executor or facade_multi" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${T}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $G --steps 20 --warmup 5 > $O/${T}_bench_n$G.json 2> $O/${T}_bench_n$G.err; echo "bench N=$G rc=$?"; summ bench_n$G $O/${T}_bench_n$G.json; tail -3 $O/${T}_bench_n$G.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 --impl reference > $O/${T}_ref_n$G.json 2> $O/${T}_ref_n$G.err; echo "ref N=$G rc=$?"; tail -c 300 $O/${T}_ref_n$G.json
python tests/multi_bench.py $G > $O/${T}_multi_exec_n$G.json 2> $O/${T}_multi_exec_n$G.err; echo "multi_bench rc=$?"; cat $O/${T}_multi_exec_n$G.json; tail -2 $O/${T}_multi_exec_n$G.err
