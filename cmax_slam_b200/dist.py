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
    """Event-sharded contrast functor: rank r holds the events of its time slab and all knots.  Two exchanges:
      mode "plane": begin (local scatter) -> all-reduce SUM of the un-blurred IL plane -> end (blur, variance, adjoint --
                    replicated on every rank --, gather over own events) -> all-reduce SUM of the partial gradients;
      mode "bands": the image phases are sharded too: reduce-scatter of the IL by row band (+ halo) -> blur of the band ->
                    all-reduce of (S1, S2) -> adjoint blur of the band -> all-gather of G -> gather -> gradient all-reduce.
                    Same bytes over NVLink, 1/world of the blur / adjoint work per rank (cmaxb_be_shard_*).
      mode "p2p":   no collective call on the data path: the kernels exchange over peer memory (CUDA IPC, NVLink), and only
                    the panorama tiles a rank's events touched travel (cmaxb_be_exchange_* / cmaxb_be_xeval, csrc/be_xchg.cuh).
                    Needs connect() once (collective) and all ranks on one node.
    Results are identical on all ranks.  "auto" = bands when the bands are thick enough (>= 2 r + 1 rows), else plane."""

    def __init__(self, warper, group=None, mode="auto"):
        self.w = warper
        self.group = group
        self.mode = mode
        self._first = True

    def set_window(self, events, knots_xyzw, t0_ns, dt_ns, n_fixed, t_next_win_beg, IGp=None, alpha=float("nan"),
                   batch_size=100):
        import math
        import torch.distributed as dist
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        beg, end = time_slab(len(events), batch_size, rank, world)
        self.slab = (beg, end)
        self.w.set_window(events[beg:end], knots_xyzw, t0_ns, dt_ns, n_fixed, t_next_win_beg, IGp, alpha)
        self._alpha_pending = isinstance(alpha, float) and math.isnan(alpha)

    def connect(self):
        """mode "p2p": open the peers' exchange blocks (collective)."""
        self.w.exchange_connect(self.group)
        self._connected = True

    def _use_bands(self, world):
        if self.mode in ("plane", "p2p") or world < 2 or self._alpha_pending:
            return False      # the window's first evaluation fixes alpha from the WHOLE summed IL: plane path
        r = self.w.blur_radius
        thick = -(-self.w.pano_height // world) >= 2 * r + 1 and (world - 1) * -(-self.w.pano_height // world) < self.w.pano_height
        if self.mode == "bands" and not thick:
            raise ValueError("bands thinner than the blur halo: use mode='plane'")
        return thick

    def eval(self, x=None, want_grad=True):
        import torch.distributed as dist
        multi = dist.is_initialized() and dist.get_world_size(self.group) > 1
        if not multi:
            self.w.eval_begin(x, want_grad)
            return self.w.eval_end()
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.mode == "p2p" and not self._alpha_pending:
            if not getattr(self, "_connected", False):
                raise RuntimeError("mode 'p2p': call connect() first")
            return self.w.xeval(x, want_grad)
        if self._use_bands(world):
            return self._eval_bands(x, want_grad, world, rank)
        self.w.eval_begin(x, want_grad)
        plane = self.w.il_plane_tensor()
        dist.all_reduce(plane, group=self.group)          # the path's one real exchange step
        self._alpha_pending = False
        # blur / contrast / adjoint / gather queued; the partial gradients are summed on the device (no host hop)
        self.w.eval_end_launch()
        if want_grad:
            gt = self.w.grad_tensor()
            if gt is not None:
                dist.all_reduce(gt, group=self.group)
        return self.w.eval_end_fetch()

    def _eval_bands(self, x, want_grad, world, rank):
        import torch
        import torch.distributed as dist
        send, recv = self.w.shard_begin(x, want_grad, world, rank)
        if dist.get_backend(self.group) == "gloo":       # CPU-side test backend: no reduce_scatter_tensor / CUDA tensors
            full = send.clone()
            dist.all_reduce(full, group=self.group)
            recv.copy_(full.view(world, -1)[rank])
        else:
            dist.reduce_scatter_tensor(recv, send, group=self.group)
        sums = self.w.shard_image()
        dist.all_reduce(sums, group=self.group)
        g_own, g_full = self.w.shard_adjoint(world)
        if g_own is not None:
            if dist.get_backend(self.group) == "gloo":
                g_full.zero_()
                g_full.view(world, -1)[rank].copy_(g_own)
                dist.all_reduce(g_full, group=self.group)
            else:
                dist.all_gather_into_tensor(g_full, g_own, group=self.group)
            self.w.shard_gather()
            gt = self.w.grad_tensor()
            if gt is not None:
                dist.all_reduce(gt, group=self.group)
        return self.w.eval_end_fetch()
