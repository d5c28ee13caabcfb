#!/bin/bash
# round 2, call p (2 GPUs): peer-memory exchange after the dirty-bitmap fix -- tests, then wall + kernel-class times of the three exchanges
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py -x -q 2>&1 | tail -3
for c in C4 C5; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 scratch/be_p2p_probe.py $c 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
done
