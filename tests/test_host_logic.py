"""Host-side logic: synthetic workloads, GSL-callback sign conventions, hypothesis sharding
(world_size-2 gloo), and the product's SO(3) math compiled for the host against the real basalt."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_deterministic_and_sorted():
    from cmax_slam_b200 import synth
    a = synth.fe_config("C1", 0.05)
    b = synth.fe_config("C1", 0.05)
    assert np.array_equal(a.events, b.events)
    t = a.events["sec"].astype(np.int64) * 10**9 + a.events["nsec"]
    assert np.all(np.diff(t) >= 0)
    assert np.all(a.events["nsec"] % 1000 == 0)            # integer microseconds
    assert a.events["x"].max() < a.width and a.events["y"].max() < a.height
    assert len(a.events) == 5000
    w = synth.make_be_window(3000, 8, 128, 64, 5, order=4, n_landmarks=200, n_fixed=3)
    t = w.events["sec"].astype(np.int64) * 10**9 + w.events["nsec"]
    assert t.min() >= w.t0_ns and t.max() < w.t0_ns + (8 - 4 + 1) * w.dt_ns
    assert np.allclose(np.linalg.norm(w.knots_xyzw, axis=1), 1.0)


def test_gsl_callback_sign_convention():
    """local_contrast_fdf returns -contrast / -gradient; f-only path passes want_grad=False
    (local_optim_contrast_gsl.cpp:32,40,48-54; global_optim_contrast_gsl_analytical.cpp:56-66)."""
    from cmax_slam_b200 import backend, frontend

    class Fake:
        def __init__(self):
            self.calls = []

        def eval(self, v, want_grad=True):
            self.calls.append(want_grad)
            return 2.5, (np.array([1.0, -2.0, 3.0]) if want_grad else None)

    f = Fake()
    c, g = frontend.local_contrast_fdf([0, 0, 0], f)
    assert c == -2.5 and np.array_equal(g, [-1.0, 2.0, -3.0])
    assert frontend.local_contrast_f([0, 0, 0], f) == -2.5 and f.calls[-1] is False
    assert np.array_equal(frontend.local_contrast_df([0, 0, 0], f), [-1.0, 2.0, -3.0])
    c, g = backend.global_contrast_fdf([0, 0, 0], f)
    assert c == -2.5 and np.array_equal(g, [-1.0, 2.0, -3.0])
    assert backend.global_contrast_f([0, 0, 0], f) == -2.5 and f.calls[-1] is False


def test_shard_indices_cover_everything_once():
    from cmax_slam_b200.dist import shard_indices
    for k in (0, 1, 7, 256):
        for world in (1, 2, 8):
            got = np.sort(np.concatenate([shard_indices(k, r, world) for r in range(world)]))
            assert np.array_equal(got, np.arange(k))


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from cmax_slam_b200.dist import sharded_eval
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
oms = np.arange(21, dtype=np.float64).reshape(7, 3) * 0.1
seen = []
def evaluate(o, want_grad):
    seen.append(len(o))
    return (o ** 2).sum(1), 2 * o
c, g = sharded_eval(evaluate, oms, True)
assert np.allclose(c, (oms ** 2).sum(1)) and np.allclose(g, 2 * oms), (c, g)
assert seen == [4 if int(sys.argv[3]) == 0 else 3]
c2, g2 = sharded_eval(evaluate, oms, False)
assert g2 is None and np.allclose(c2, c)
dist.destroy_process_group()
print("ok")
'''


def test_hypothesis_sharding_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)


# ---- the PRODUCT's so3_math.cuh compiled for the host, against the real basalt golden vectors ------
_SHIM = r'''
#include "so3_math.cuh"
using namespace cmaxb;
extern "C" void so3h_exp(const double* w, double* q) { Vec3 v{w[0],w[1],w[2]}; Quat r = so3_exp(v); q[0]=r.x;q[1]=r.y;q[2]=r.z;q[3]=r.w; }
extern "C" void so3h_log(const double* q, double* w) { Quat a{q[0],q[1],q[2],q[3]}; Vec3 r = so3_log(a); w[0]=r.x;w[1]=r.y;w[2]=r.z; }
extern "C" void so3h_jac(const double* w, double* Jl, double* Jli) { Vec3 v{w[0],w[1],w[2]}; Mat3 a = so3_left_jacobian(v), b = so3_left_jacobian_inv(v); for (int i=0;i<9;++i){Jl[i]=a.m[i];Jli[i]=b.m[i];} }
extern "C" void so3h_spline(int order, const double* knots, int s, double u, double* q, double* R, double* J) {
  const Quat* k = reinterpret_cast<const Quat*>(knots);
  Mat3 Jm[4]; Quat r;
  if (order == 2) r = so3_spline_eval<2>(k, s, u, Jm); else r = so3_spline_eval<4>(k, s, u, Jm);
  q[0]=r.x;q[1]=r.y;q[2]=r.z;q[3]=r.w;
  Mat3 m = quat_to_mat(r); for (int i=0;i<9;++i) R[i]=m.m[i];
  for (int b=0;b<order;++b) for (int i=0;i<9;++i) J[9*b+i]=Jm[b].m[i];
}
'''


@pytest.fixture(scope="module")
def so3h(tmp_path_factory):
    d = tmp_path_factory.mktemp("so3h")
    src = d / "shim.cpp"
    src.write_text(_SHIM)
    out = d / "libso3h.so"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I", os.path.join(ROOT, "cmax_slam_b200", "csrc"), "-o", str(out), str(src)])
    return C.CDLL(str(out))


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("order", [2, 4])
def test_product_spline_math_vs_real_basalt(so3h, golden, order):
    g = golden("spline_ref.npz")
    knots = np.ascontiguousarray(g[f"knots_{order}"])
    t0, dt = int(g["t0"]), int(g["dt"])
    so3h.so3h_spline.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double] + [C.POINTER(C.c_double)] * 3
    for i, t in enumerate(g[f"t_{order}"]):
        st = int(t) - t0
        s, u = st // dt, float(st % dt) / float(dt)
        assert s == g[f"idx_{order}"][i]
        q, R, J = np.zeros(4), np.zeros(9), np.zeros(9 * order)
        so3h.so3h_spline(order, _d(knots), int(s), u, _d(q), _d(R), _d(J))
        assert np.abs(q - g[f"q_{order}"][i]).max() < 5e-15
        assert np.abs(R.reshape(3, 3) - g[f"R_{order}"][i]).max() < 5e-15
        assert np.abs(J.reshape(order, 3, 3) - g[f"J_{order}"][i]).max() < 1e-14


def test_product_exp_log_jacobians_vs_real_sophus(so3h, golden):
    g = golden("spline_ref.npz")
    for i, w in enumerate(g["w"]):
        w = np.ascontiguousarray(w)
        q, lg, a, b = np.zeros(4), np.zeros(3), np.zeros(9), np.zeros(9)
        so3h.so3h_exp(_d(w), _d(q))
        assert np.abs(q - g["exp_w"][i]).max() < 1e-15
        so3h.so3h_log(_d(np.ascontiguousarray(g["exp_w"][i])), _d(lg))
        assert np.abs(lg - g["log_exp_w"][i]).max() < 1e-15
        so3h.so3h_jac(_d(w), _d(a), _d(b))
        assert np.abs(a - g["Jl"][i]).max() < 1e-14 and np.abs(b - g["Jlinv"][i]).max() < 1e-14


def test_time_slabs_are_batch_aligned_and_cover_the_window():
    """Event sharding of one back-end window: contiguous slabs cut at batch boundaries."""
    from cmax_slam_b200.dist import time_slab
    for n, bs, world in [(1001, 100, 2), (1000, 100, 3), (50, 100, 4), (0, 100, 2), (12345, 64, 8), (401, 100, 2)]:
        sl = [time_slab(n, bs, r, world) for r in range(world)]
        assert sl[0][0] == 0 and sl[-1][1] == n
        assert all(sl[i][1] == sl[i + 1][0] for i in range(world - 1))
        assert all(b % bs == 0 or b == n for b, _ in sl)
        # the reference's trailing single-event batch can only ever sit in the LAST slab
        for b, e in sl[:-1]:
            assert (e - b) % bs == 0 or e == n      # a ragged slab is always the global tail


# ---- time-sharded back-end window: the orchestration of ShardedEventWarper over gloo, world 2, with a numpy stand-in for the
# device warper (the kernels themselves are covered by tests/test_gpu_sharded.py and tests/test_gpu_multi.py)
_SHARD_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200.dist import ShardedEventWarper
rank = int(sys.argv[3])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=2)

class FakeWarper:
    """IL = votes of the slab's events at pixel (x + round(10 * p0)) mod W; contrast = variance of the (summed) IL;
    partial gradient = sum over OWN events of IL[pixel] * (x, y, 1) -- needs the SUMMED IL, like the adjoint gather."""
    pano_width, pano_height, blur_radius = 32, 16, 4
    def set_window(self, events, knots, t0, dt, n_fixed, tnext, IGp, alpha):
        self.ev = events; self.n_params = 3
    def _pix(self, x):
        return (self.ev["y"] % 16) * 32 + (self.ev["x"] + int(round(10 * x[0]))) % 32
    def eval_begin(self, x, want_grad):
        self.x = np.asarray(x, float); self.want = want_grad
        self.plane = torch.zeros(32 * 16, dtype=torch.float32)
        np.add.at(self.plane.numpy(), self._pix(self.x), 1.0)
    def il_plane_tensor(self):
        return self.plane
    def eval_end_launch(self):
        il = self.plane.numpy().astype(np.float64)
        self.c = float(il.var())
        w = il[self._pix(self.x)]
        self.g = torch.tensor([float((w * self.ev["x"]).sum()), float((w * self.ev["y"]).sum()), float(w.sum())], dtype=torch.float64)
    def grad_tensor(self):
        return self.g
    def eval_end_fetch(self):
        return self.c, (self.g.numpy().copy() if self.want else None)
    def eval_end(self):
        self.eval_end_launch(); return self.eval_end_fetch()

rng = np.random.default_rng(5)
ev = np.zeros(1001, dtype=[("x", np.int64), ("y", np.int64)])
ev["x"] = rng.integers(0, 32, 1001); ev["y"] = rng.integers(0, 16, 1001)
sh = ShardedEventWarper(FakeWarper(), mode="plane")
sh.set_window(ev, None, 0, 0, 1, None, None, 0.5, batch_size=100)
x = np.array([0.3, 0.0, 0.0])
c, g = sh.eval(x, True)
c_v, g_v = sh.eval(x, False)
one = FakeWarper(); one.set_window(ev, None, 0, 0, 1, None, None, 0.5); one.eval_begin(x, True); c1, g1 = one.eval_end()
assert sh.slab == ((0, 600) if rank == 0 else (600, 1001)), sh.slab      # 11 batches of 100: 6 + 5
assert abs(c - c1) < 1e-12 and abs(c_v - c1) < 1e-12 and g_v is None
assert np.allclose(g, g1, rtol=1e-12), (g, g1)
try:
    sh.mode = "p2p"; sh.eval(x, True)
    raise SystemExit("p2p without connect() must be refused")
except RuntimeError:
    pass
sh.mode = "bands"                              # 16 rows / 2 ranks = 8 < 2 r + 1 = 9: refused
try:
    sh._use_bands(2)
    raise SystemExit("bands thinner than the halo must be refused")
except ValueError:
    pass
dist.barrier(); dist.destroy_process_group()
print("ok")
'''


def test_time_sharded_orchestration_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_SHARD_WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs)
