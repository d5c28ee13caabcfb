#!/bin/bash
# round 2, call q (8 GPUs): wall + kernel-class times of the three back-end exchanges, C5 and C4 sharded by time over 8
mkdir -p gpurun_out
for c in C5 C4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 scratch/be_p2p_probe.py $c 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -12
done
