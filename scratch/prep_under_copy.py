# Does a concurrent H2D copy slow the packet preparation / the evaluation kernel?  (e2e timeline: preparation 62 us with a copy in flight, 30 us without)
import os, sys, time, ctypes as C, threading
sys.path.insert(0, '.')
import numpy as np, torch
from cmax_slam_b200 import synth, _capi
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pkt = synth.fe_config("C2")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
fe = AngVelEstimatorCMax(pkt.width, pkt.height, pkt.K, pkt.lut, device=0, stream=stream.cuda_stream, lanes=1, packet_slots=2)
n = len(pkt.events)
dev_ev = torch.empty(16 * n, dtype=torch.uint8, device="cuda")
dev_ev.copy_(torch.from_numpy(pkt.events.view(np.uint8).reshape(-1).copy()))
L = _capi.lib(); FE = fe._h
om = (C.c_double * 3)(0.3, -0.2, 0.5); rc_ = (C.c_double * 1)(); rg = (C.c_double * 3)()
ptr = C.c_void_p(dev_ev.data_ptr())
def prep():
    L.cmaxb_fe_set_packet_view(FE, ptr, n, float(pkt.t_ref_sec))
def ev():
    L.cmaxb_fe_eval_launch(FE, om, 1, 1); L.cmaxb_fe_eval_fetch(FE, rc_, rg)
prep(); ev(); torch.cuda.synchronize()
copy_stream = torch.cuda.Stream()
src = torch.empty(3 << 20, dtype=torch.uint8).pin_memory(); src.numpy()[:] = 3
dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
d2d_src = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps, bg=None):
    stop = False
    if bg:
        # keep ~40 background copies queued ahead on the copy stream for the whole measurement
        with torch.cuda.stream(copy_stream):
            for i in range(400): bg(i)
    time.sleep(0.0005)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream)
    e1.synchronize()
    busy = not copy_stream.query()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, busy
h2d = lambda i: dst[(i % 16) * (3 << 20):(i % 16 + 1) * (3 << 20)].copy_(src, non_blocking=True)
d2d = lambda i: dst[:3 << 20].copy_(d2d_src[:3 << 20], non_blocking=True)
# SM-driven pull: the pinned buffer seen as a device tensor (zero-copy mapping), copied by an elementwise kernel instead of the DMA engine
import ctypes as _C
_rt = _C.CDLL("libcudart.so")
_dp = _C.c_void_p()
assert _rt.cudaHostGetDevicePointer(_C.byref(_dp), _C.c_void_p(src.data_ptr()), 0) == 0
class _V:
    __cuda_array_interface__ = {"shape": (3 << 20,), "typestr": "|u1", "data": (_dp.value, False), "version": 2}
src_dev_view = torch.as_tensor(_V(), device="cuda")
pull = lambda i: dst[(i % 16) * (3 << 20):(i % 16 + 1) * (3 << 20)].copy_(src_dev_view, non_blocking=True)
d2h_buf = torch.empty(3 << 20, dtype=torch.uint8).pin_memory()
d2h = lambda i: d2h_buf.copy_(dst[:3 << 20], non_blocking=True)
for name, fn, reps in (("prep", prep, 100), ("eval f+g (sync each)", ev, 100)):
    for bname, bg in (("alone", None), ("with H2D 3 MB copies", h2d), ("with SM-pulled 3 MB copies", pull), ("with D2D 3 MB copies", d2d), ("with D2H 3 MB copies", d2h), ("alone again", None)):
        us, busy = timed(fn, reps, bg)
        print("%-22s %-24s %7.1f us   (copy stream still busy at the end: %s)" % (name, bname, us, busy))

torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(copy_stream):
    for i in range(5): pull(i)
    e0.record(copy_stream)
    for i in range(50): pull(i)
    e1.record(copy_stream)
torch.cuda.synchronize()
print("SM-pulled 3 MB copy alone: %.1f us (%.1f GB/s)" % (e0.elapsed_time(e1) / 50 * 1e3, (3 << 20) / (e0.elapsed_time(e1) / 50 * 1e-3) / 1e9))
with torch.cuda.stream(copy_stream):
    for i in range(5): h2d(i)
    e0.record(copy_stream)
    for i in range(50): h2d(i)
    e1.record(copy_stream)
torch.cuda.synchronize()
print("DMA 3 MB copy alone: %.1f us (%.1f GB/s)" % (e0.elapsed_time(e1) / 50 * 1e3, (3 << 20) / (e0.elapsed_time(e1) / 50 * 1e-3) / 1e9))
