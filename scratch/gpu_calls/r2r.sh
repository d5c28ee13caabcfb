#!/bin/bash
# round 2, call r (8 GPUs): peer-memory exchange after the bitmap / staging / reduction changes -- tests, then the exchange probe on C5 and C4
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_be.py -x -q 2>&1 | tail -3
for c in C5 C4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scratch/be_p2p_probe.py $c 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -13
done
