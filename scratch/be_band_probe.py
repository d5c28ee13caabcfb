# 1-GPU probe: per-kernel times of one rank's share of a time-sharded back-end window (rank 0 of `world`), whole-plane path
# (quad gather, full-panorama blur / adjoint) against the row-band path's pieces (pack, band blur, band adjoint, plane gather)
import sys, json
sys.path.insert(0, '.')
import numpy as np, torch
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
from cmax_slam_b200.dist import time_slab
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
w = synth.be_config(name, device="cuda")
rng = np.random.default_rng(5)
IGp = np.abs(rng.normal(0, 0.3, (w.pano_height, w.pano_width))).astype(np.float32)
x = rng.normal(0, 0.01, 3 * (len(w.knots_xyzw) - w.n_fixed))
b, e = time_slab(len(w.events), 100, 0, world)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
be = EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, w.pano_width, w.pano_height, spline_order=2, stream=st.cuda_stream)
be.set_window(w.events[b:e], w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
for _ in range(3): be.eval(x, True)
be.profile(True)
for _ in range(5): be.eval(x, True)
kt = be.kernel_times()
print(name, "rank 0 of", world, "events", e - b, "PLAIN (quad gather, full image):", {k: round(v[0] / v[1] * 1e3, 1) for k, v in kt.items()})
be.profile(False)
# the band path's compute pieces with rank 0's band geometry of `world` ranks (collectives left out: buffers used as they are)
def band_eval(grad):
    send, recv = be.shard_begin(x, grad, world, 0)
    recv.copy_(send.view(world, -1)[0])
    be.shard_image()
    own, full = be.shard_adjoint(world)
    if own is not None:
        be.shard_gather()
    return be.eval_end_fetch()
for _ in range(3): band_eval(True)
be.profile(True)
for _ in range(5): band_eval(True)
kt = be.kernel_times()
print(name, "BANDS compute pieces:", {k: round(v[0] / v[1] * 1e3, 1) for k, v in kt.items()})
be.profile(False)
import time
for fn, lab in ((lambda: be.eval(x, True), "plain f+g"), (lambda: band_eval(True), "bands f+g (no collectives)"), (lambda: be.eval(x, False), "plain value"), (lambda: band_eval(False), "bands value")):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); print(lab, round((time.perf_counter() - t) / 10 * 1e6, 1), "us")
