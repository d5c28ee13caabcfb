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


def sharded_eval_fused(fe, omegas, want_grad=True, group=None):
    """Hypothesis sharding with the result exchange fused into the evaluation kernel (BASELINE config C3): `fe` is an
    AngVelEstimatorCMax on which exchange_connect(group) has been called on every rank.  Rank r evaluates hypotheses
    r, r + N, ... in ONE launch whose publishing CTA also stores the rows into every peer's buffer over NVLink and
    collects the peers' rows -- no collective call here.  K must be a multiple of the world size with K / N <= 32.
    Returns (contrasts[K], grads[K,3] or None), identical on every rank."""
    import torch.distributed as dist
    om = np.ascontiguousarray(omegas, dtype=np.float64).reshape(-1, 3)
    K = om.shape[0]
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if K % world or K // world > 32 or K // world > fe.max_hypotheses:
        raise ValueError("K must be a multiple of the world size with K / world <= min(32, max_hypotheses)")
    fe.eval_launch(om[shard_indices(K, rank, world)], want_grad)
    rows = fe.eval_fetch_all()                       # [world, K / world, 4]
    out = np.zeros((K, 4))
    for r in range(world):
        out[shard_indices(K, r, world)] = rows[r]
    return out[:, 0].copy(), (out[:, 1:].copy() if want_grad else None)


# ------------------------------------------------------------------------------------------------
# One big back-end window sharded by TIME across ranks (BASELINE config C5)
# ------------------------------------------------------------------------------------------------
def time_slab(n_events, batch_size, rank, world):
    """[begin, end) of the events owned by `rank`.  Slabs are contiguous in time (events are time-sorted) and
    cut at multiples of batch_size, so that every rank forms exactly the batches (and batch mid-times) of the
    un-sharded evaluation; the tail -- including the reference's never-visited trailing single event
    (event_pano_warper.cpp:188-196) -- stays with the last rank."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    nb = (n_events + batch_size - 1) // batch_size
    per = (nb + world - 1) // world
    b0 = min(rank * per, nb)
    b1 = min((rank + 1) * per, nb)
    beg = min(b0 * batch_size, n_events)
    end = n_events if rank == world - 1 else min(b1 * batch_size, n_events)
    if rank == world - 1 and b0 >= nb:
        beg = n_events
    return beg, max(beg, end)


class ShardedEventWarper:
    """Event-sharded contrast functor: rank r holds the events of its time slab and all knots.
    eval(x) = begin (local scatter) -> all-reduce SUM of the un-blurred IL plane (image sized, NCCL over NVLink)
    -> end (blur, variance -- identical on every rank --, adjoint, gather over own events) -> all-reduce SUM
    of the partial gradients (3*K_opt doubles).  Results are identical on all ranks."""

    def __init__(self, warper, group=None):
        self.w = warper
        self.group = group

    def set_window(self, events, knots_xyzw, t0_ns, dt_ns, n_fixed, t_next_win_beg, IGp=None, alpha=float("nan"),
                   batch_size=100):
        import torch.distributed as dist
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        beg, end = time_slab(len(events), batch_size, rank, world)
        self.slab = (beg, end)
        self.w.set_window(events[beg:end], knots_xyzw, t0_ns, dt_ns, n_fixed, t_next_win_beg, IGp, alpha)

    def eval(self, x=None, want_grad=True):
        import torch
        import torch.distributed as dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        self.w.eval_begin(x, want_grad)
        if multi:
            plane = self.w.il_plane_tensor()
            dist.all_reduce(plane, group=self.group)          # the path's one real exchange step
        if not multi:
            return self.w.eval_end()
        # blur / contrast / adjoint / gather queued; the partial gradients are summed on the device (no host hop)
        self.w.eval_end_launch()
        if want_grad:
            gt = self.w.grad_tensor()
            if gt is not None:
                dist.all_reduce(gt, group=self.group)
        return self.w.eval_end_fetch()
