#!/bin/bash
# developer tool (GPU box, 8 GPUs): PCIe ceiling with all GPUs copying, bench.py at N = 8, the multi-GPU executor on 8 devices
T=${1:-r02m}; G=${2:-8}; O=gpurun_out; mkdir -p $O
. tests/gpu_summ.sh
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1; nproc > $O/${T}_nproc.txt
python tests/pcie_ceiling.py $G > $O/${T}_pcie_ceiling.json 2> $O/${T}_pcie_ceiling.err; echo "pcie rc=$?"; cat $O/${T}_pcie_ceiling.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $G --steps 20 --warmup 5 > $O/${T}_bench_n$G.json 2> $O/${T}_bench_n$G.err; echo "bench N=$G rc=$?"; summ bench_n$G $O/${T}_bench_n$G.json; grep -v "^$" $O/${T}_bench_n$G.err | grep -A30 "Fatal Python\|Segmentation\|Error" | head -40
python tests/multi_bench.py $G > $O/${T}_multi_exec_n$G.json 2> $O/${T}_multi_exec_n$G.err; echo "multi_bench rc=$?"; cat $O/${T}_multi_exec_n$G.json; tail -2 $O/${T}_multi_exec_n$G.err
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k "multi_executor_two_gpus" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${T}_pytest.log
