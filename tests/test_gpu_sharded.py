"""Event-sharded back-end evaluation (one window split by time): begin / IL-plane exchange /
end.  On one GPU: (a) begin+end with no exchange == eval == oracle; (b) two processes sharing cuda:0, each with
a time slab, exchanging the IL plane and the partial gradients with gloo (NCCL refuses two ranks on one
device) == the un-sharded evaluation."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cmax_slam_b200 import synth
from cmax_slam_b200._capi import CmaxbError

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K_T = (60.0, 61.0, 31.5, 23.5)


def _window(n=20001, order=2):
    return synth.make_be_window(n, 8, 128, 64, 41, order=order, sensor=(64, 48), K4=K_T, n_landmarks=300,
                                n_fixed=1 if order == 2 else 3)


@pytest.mark.parametrize("order", [2, 4])
def test_begin_end_equals_eval_and_oracle(oracle, order):
    from cmax_slam_b200.backend import EventWarperCMax
    w = _window(order=order)
    rng = np.random.default_rng(3)
    IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
    be = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=order)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    x = rng.normal(0, 0.02, 3 * (8 - w.n_fixed))
    c0, g0 = be.eval(x, True)
    be.eval_begin(x, True)
    plane = be.il_plane_tensor()
    ilo, iln = be.local_iwe(x)
    be.eval_begin(x, True)
    assert np.abs(plane.cpu().numpy().reshape(64, 128) - (ilo + iln)).max() <= 4e-6 * max(1.0, float((ilo + iln).max()))
    c1, g1 = be.eval_end()
    assert abs(c1 - c0) <= 1e-7 * abs(c0) and np.abs(g1 - g0).max() <= 1e-6 * np.abs(g0).max()
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, order, w.n_fixed, w.tnext, IGp, 0.5)
    ro = oracle.be_eval(a, x, True)
    assert abs(c1 - ro["contrast"]) <= 1e-5 * ro["contrast"]
    assert np.abs(g1 - ro["grad"]).max() <= 1e-5 * np.abs(ro["grad"]).max()
    be.eval_begin(x, False)
    c2, g2 = be.eval_end()
    assert g2 is None and abs(c2 - c0) <= 1e-7 * abs(c0)
    be.close()


_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
from cmax_slam_b200.dist import ShardedEventWarper
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=2)
torch.cuda.set_device(0)
K_T = (60.0, 61.0, 31.5, 23.5)
w = synth.make_be_window(20001, 8, 128, 64, 41, order=2, sensor=(64, 48), K4=K_T, n_landmarks=300, n_fixed=1)
rng = np.random.default_rng(3)
IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
x = rng.normal(0, 0.02, 21)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
sh = ShardedEventWarper(EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2, stream=stream.cuda_stream))
sh.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, float("nan"))
c, g = sh.eval(x, True)
c_f, _ = sh.eval(x, False)          # alpha fixed by now: row-band path, value only
c_b, g_b = sh.eval(x, True)         # row-band path with gradient
assert sh._use_bands(2)
alpha = sh.w.alpha
np.save(sys.argv[4] + f"/r{rank}.npy", np.concatenate([[c, c_f, alpha, sh.slab[0], sh.slab[1]], g]))
np.save(sys.argv[4] + f"/b{rank}.npy", np.concatenate([[c_b], g_b]))
# exchange by the kernels over peer memory (CUDA IPC between the two processes), sparse by dirty tile
sh.mode = "p2p"
sh.connect()
outs = []
for i, (xx, wg) in enumerate([(x, True), (x, False), (0.5 * x, True), (x, True)]):
    c_p, g_p = sh.eval(xx, wg)
    outs.append(np.concatenate([[c_p], g_p if wg else np.zeros(21)]))
c_plain, g_plain = sh.w.eval(x, True)      # a plain (own-slab) evaluation in between breaks the tile invariant on purpose
c_p, g_p = sh.eval(x, True)
outs.append(np.concatenate([[c_p], g_p]))
np.save(sys.argv[4] + f"/p{rank}.npy", np.stack(outs))
sh.w.exchange_close()
dist.barrier()
dist.destroy_process_group()
print("ok")
'''


def test_two_ranks_time_sharded_equals_unsharded(oracle, tmp_path):
    import socket
    from cmax_slam_b200.backend import EventWarperCMax
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0[:3], r1[:3]) or np.allclose(r0[:3], r1[:3], rtol=1e-12)     # identical on both ranks
    assert np.allclose(r0[5:], r1[5:], rtol=1e-12)
    assert (r0[3], r0[4], r1[3], r1[4]) == (0, 10100, 10100, 20001)                      # batch-aligned slabs
    # un-sharded reference on this process
    w = _window()
    rng = np.random.default_rng(3)
    IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
    x = rng.normal(0, 0.02, 21)
    be = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, float("nan"))
    c, g = be.eval(x, True)
    assert abs(be.alpha - r0[2]) <= 1e-6 * be.alpha                                       # alpha from the SUMMED IL
    assert abs(r0[0] - c) <= 1e-6 * abs(c) and abs(r0[1] - c) <= 1e-6 * abs(c)
    assert np.abs(r0[5:] - g).max() <= 1e-6 * np.abs(g).max()
    a = oracle.be_args(w.events, w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext, IGp, be.alpha)
    ro = oracle.be_eval(a, x, True)
    assert abs(r0[0] - ro["contrast"]) <= 1e-5 * ro["contrast"]
    assert np.abs(r0[5:] - ro["grad"]).max() <= 1e-5 * np.abs(ro["grad"]).max()
    # the row-band exchange (reduce-scatter / all-gather) gives the same numbers on both ranks
    b0, b1 = np.load(tmp_path / "b0.npy"), np.load(tmp_path / "b1.npy")
    assert np.allclose(b0, b1, rtol=1e-12)
    assert abs(b0[0] - c) <= 1e-6 * abs(c) and np.abs(b0[1:] - g).max() <= 1e-6 * np.abs(g).max()
    # the peer-memory exchange: identical on both ranks bit for bit (sums formed in rank order), equal to the un-sharded numbers
    p0, p1 = np.load(tmp_path / "p0.npy"), np.load(tmp_path / "p1.npy")
    assert np.array_equal(p0, p1)
    c_h, g_h = be.eval(0.5 * x, True)
    for i in (0, 3, 4):
        assert abs(p0[i, 0] - c) <= 1e-6 * abs(c) and np.abs(p0[i, 1:] - g).max() <= 1e-6 * np.abs(g).max()
    assert abs(p0[1, 0] - c) <= 1e-6 * abs(c)
    assert abs(p0[2, 0] - c_h) <= 1e-6 * abs(c_h) and np.abs(p0[2, 1:] - g_h).max() <= 1e-6 * np.abs(g_h).max()
    be.close()


@pytest.mark.parametrize("world,pano", [(2, (128, 64)), (3, (128, 64)), (4, (256, 100))])
def test_row_band_sharding_emulated_in_one_process(oracle, world, pano):
    """cmaxb_be_shard_*: `world` handles on one GPU play the ranks; the collectives are emulated with torch ops on the
    library's own buffers.  Uneven bands (64 rows / 3 ranks, 100 rows / 4), halo rows off the panorama, IGp, value-only
    and gradient evaluations -- against the un-sharded evaluation and the oracle."""
    import torch
    from cmax_slam_b200.backend import EventWarperCMax
    from cmax_slam_b200.dist import time_slab
    pw, ph = pano
    w = synth.make_be_window(24001, 8, pw, ph, 77, order=2, sensor=(64, 48), K4=K_T, n_landmarks=300, n_fixed=1)
    rng = np.random.default_rng(5)
    IGp = np.abs(rng.normal(0, 0.3, (ph, pw))).astype(np.float32)
    x = rng.normal(0, 0.02, 21)
    st = torch.cuda.Stream()                 # the handles work on torch's stream: library launches and torch ops stay ordered
    torch.cuda.set_stream(st)
    ref = EventWarperCMax(64, 48, w.lut, pw, ph, spline_order=2, stream=st.cuda_stream)
    ref.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    c0, g0 = ref.eval(x, True)
    hs = []
    for r in range(world):
        b, e = time_slab(len(w.events), 100, r, world)
        h = EventWarperCMax(64, 48, w.lut, pw, ph, spline_order=2, stream=st.cuda_stream)
        h.set_window(w.events[b:e], w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
        hs.append(h)
    for want_grad in (True, False, True):
        bufs = [h.shard_begin(x, want_grad, world, r) for r, h in enumerate(hs)]
        total = sum(b[0] for b in bufs)                                  # reduce ...
        for r, (_, recv) in enumerate(bufs):
            recv.copy_(total.view(world, -1)[r])                         # ... scatter
        sums = [h.shard_image() for h in hs]
        tot = sum(sums)
        for s_ in sums:
            s_.copy_(tot)
        outs = [h.shard_adjoint(world) for h in hs]
        if want_grad:
            full = torch.cat([o[0] for o in outs])                       # all-gather
            for h, o in zip(hs, outs):
                o[1].copy_(full)
                h.shard_gather()
            gts = [h.grad_tensor() for h in hs]
            gsum = sum(gts)
            for gt in gts:
                gt.copy_(gsum)
        res = [h.eval_end_fetch() for h in hs]
        for c, g in res:
            assert abs(c - c0) <= 1e-6 * abs(c0)
            if want_grad:
                assert np.abs(g - g0).max() <= 1e-6 * np.abs(g0).max()
            else:
                assert g is None
    torch.cuda.set_stream(torch.cuda.default_stream())
    a = oracle.be_args(w.events, w.lut, 64, 48, pw, ph, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext, IGp, 0.5)
    ro = oracle.be_eval(a, x, True)
    assert abs(res[0][0] - ro["contrast"]) <= 1e-5 * ro["contrast"]
    assert np.abs(res[0][1] - ro["grad"]).max() <= 1e-5 * np.abs(ro["grad"]).max()
    # thin bands are refused, and so is a window whose alpha is still to be computed
    with pytest.raises(CmaxbError):
        hs[0].shard_begin(x, True, 16, 0)
    hs[0].set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, float("nan"))
    with pytest.raises(CmaxbError):
        hs[0].shard_begin(x, True, world, 0)
    for h in hs + [ref]:
        h.close()


def test_plain_eval_after_a_sharded_evaluation(oracle):
    """A handle that ran begin/end (its IL assembled into the exchange plane) must return to the scatter's own
    accumulators for plain evaluations -- at other parameters and after a new window (round-1 advisor finding:
    il_is_plane was never reset, so the contrast came from the stale plane)."""
    from cmax_slam_b200.backend import EventWarperCMax
    w = _window()
    rng = np.random.default_rng(3)
    IGp = np.abs(rng.normal(0, 0.3, (64, 128))).astype(np.float32)
    x = rng.normal(0, 0.02, 21)
    x2 = rng.normal(0, 0.05, 21)
    be = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2)
    ref = EventWarperCMax(64, 48, w.lut, 128, 64, spline_order=2)
    for h in (be, ref):
        h.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.4)
    be.eval_begin(x, True)
    c_s, g_s = be.eval_end()
    c_r, g_r = ref.eval(x, True)
    assert abs(c_s - c_r) <= 1e-6 * abs(c_r) and np.abs(g_s - g_r).max() <= 1e-6 * np.abs(g_r).max()
    # plain evaluation at OTHER parameters on the handle that was sharded
    c2, g2 = be.eval(x2, True)
    c2r, g2r = ref.eval(x2, True)
    assert abs(c2 - c2r) <= 1e-6 * abs(c2r), (c2, c2r, c_s)      # two handles: f32 atomic order
    assert np.abs(g2 - g2r).max() <= 1e-5 * np.abs(g2r).max()
    assert np.allclose(be.computeImageOfWarpedEvents(x2), ref.computeImageOfWarpedEvents(x2), rtol=0, atol=1e-4)
    # ... and after a new window (half of the events)
    n2 = 10000
    for h in (be, ref):
        h.set_window(w.events[:n2], w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.4)
    c3, _ = be.eval(x2, True)
    c3r, _ = ref.eval(x2, True)
    assert abs(c3 - c3r) <= 1e-6 * abs(c3r)
    a = oracle.be_args(w.events[:n2], w.lut, 64, 48, 128, 64, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext, IGp, 0.4)
    assert abs(c3 - oracle.be_eval(a, x2, False)["contrast"]) <= 1e-5 * abs(c3)
    be.close(); ref.close()


@pytest.mark.parametrize("pano", [(128, 64), (256, 100), (132, 70)])
def test_peer_exchange_single_rank_equals_eval(oracle, pano):
    """cmaxb_be_xeval with a world of one rank (the rank is its own peer): dirty-tile clean / push / pull and the tagged-word
    exchanges run exactly as at N ranks; result == plain evaluation == oracle, repeatedly and at changing parameters."""
    from cmax_slam_b200.backend import EventWarperCMax
    import ctypes as C
    from cmax_slam_b200 import _capi
    pw, ph = pano
    w = synth.make_be_window(16001, 8, pw, ph, 78, order=2, sensor=(64, 48), K4=K_T, n_landmarks=300, n_fixed=1)
    rng = np.random.default_rng(6)
    IGp = np.abs(rng.normal(0, 0.3, (ph, pw))).astype(np.float32)
    be = EventWarperCMax(64, 48, w.lut, pw, ph, spline_order=2)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
    h = C.create_string_buffer(64)
    _capi.check(be._L.cmaxb_be_exchange_init(be._h, 1, 0, h))
    _capi.check(be._L.cmaxb_be_exchange_connect(be._h, h.raw))
    for k in range(4):
        x = rng.normal(0, 0.02, 21)
        c0, g0 = be.eval(x, True) if k % 2 == 0 else (None, None)     # interleaved plain evaluations (tile invariant reset)
        c1, g1 = be.xeval(x, True)
        c2, _ = be.xeval(x, False)
        if c0 is None:
            c0, g0 = be.eval(x, True)
        assert abs(c1 - c0) <= 1e-6 * abs(c0) and abs(c2 - c0) <= 1e-6 * abs(c0)
        assert np.abs(g1 - g0).max() <= 1e-6 * np.abs(g0).max()
    a = oracle.be_args(w.events, w.lut, 64, 48, pw, ph, w.knots_xyzw, w.t0_ns, w.dt_ns, 2, w.n_fixed, w.tnext, IGp, 0.5)
    ro = oracle.be_eval(a, x, True)
    assert abs(c1 - ro["contrast"]) <= 1e-5 * ro["contrast"]
    assert np.abs(g1 - ro["grad"]).max() <= 1e-5 * np.abs(ro["grad"]).max()
    be.exchange_close()
    be.close()
