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
c_f, _ = sh.eval(x, False)
alpha = sh.w.alpha
np.save(sys.argv[4] + f"/r{rank}.npy", np.concatenate([[c, c_f, alpha, sh.slab[0], sh.slab[1]], g]))
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
    be.close()


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
