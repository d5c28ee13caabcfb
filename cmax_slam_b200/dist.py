"""Multi-GPU hypothesis sharding (BASELINE config C3).

Units = (event packet x angular-velocity hypothesis) evaluations; they are independent, so they
shard across ranks with no data-path collective.  The packet is replicated (16 MB), rank r
evaluates hypotheses k with k % world == r, and the per-hypothesis rows (contrast, g0, g1, g2) are
combined with ONE all-reduce of a zero-padded [K,4] f64 buffer (NCCL over NVLink on GPUs, gloo in
the CPU tests).  One process per GPU; torch.distributed is plumbing only.
"""
import numpy as np


def shard_indices(k, rank, world):
    """Hypothesis indices owned by `rank` (round-robin: neighbouring hypotheses have similar cost)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return np.arange(rank, k, world)


def sharded_eval(evaluate, omegas, want_grad=True, group=None, device=None):
    """evaluate(omegas_local[kl,3], want_grad) -> (contrasts[kl], grads[kl,3] or None) on this rank.
    Returns (contrasts[K], grads[K,3] or None) identical on every rank."""
    import torch
    import torch.distributed as dist

    om = np.ascontiguousarray(omegas, dtype=np.float64).reshape(-1, 3)
    K = om.shape[0]
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard_indices(K, rank, world)
    rows = np.zeros((K, 4))
    if len(mine):
        c, g = evaluate(om[mine], want_grad)
        rows[mine, 0] = c
        if want_grad:
            rows[mine, 1:] = g
    if world > 1:
        buf = torch.from_numpy(rows)
        if device is not None:
            buf = buf.to(device)
        dist.all_reduce(buf, group=group)  # zero-padded rows: SUM == gather
        rows = buf.cpu().numpy()
    return rows[:, 0].copy(), (rows[:, 1:].copy() if want_grad else None)
