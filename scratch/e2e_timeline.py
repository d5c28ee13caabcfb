# Poor man's timeline of the end-to-end loop (no nsys in the image): CUDA events after every upload (copy stream), packet preparation and
# evaluation (main stream, lanes = 1 so that the evaluation runs there); prints when each finished relative to the first event.
import os, sys, time, ctypes as C
sys.path.insert(0, '.')
os.environ.setdefault("CMAXB_FE_LANES", "1")
import numpy as np, torch
from cmax_slam_b200 import synth, _capi
from cmax_slam_b200.frontend import AngVelEstimatorCMax
from cmax_slam_b200.stream import EventStream
dev = torch.device("cuda", 0)
pkt = synth.fe_config("C2")
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
fe = AngVelEstimatorCMax(pkt.width, pkt.height, pkt.K, pkt.lut, device=0, stream=stream.cuda_stream, lanes=1, packet_slots=4)
per_packet = len(pkt.events)
steps = 14
n_total = int((steps + 12) * 2.0e5 + 1.6 * per_packet)
ev, _ = synth.make_fe_stream(n_total, pkt.width, pkt.height, pkt.K, 11)
pinned = torch.empty(len(ev) * 16, dtype=torch.uint8).pin_memory()
pinned.numpy()[:] = ev.view(np.uint8).reshape(-1)
base = pinned.data_ptr()
t_ns = ev["sec"].astype(np.int64) * 1_000_000_000 + ev["nsec"]
edges = np.searchsorted(t_ns, t_ns[0] + (np.arange(1, steps + 60) * 10_000_000))
edges = np.concatenate([[0], edges[edges < len(ev)], [len(ev)]])
st = EventStream(0.01, per_packet, 1)
copy_stream = torch.cuda.Stream()
st.attach_device(0, copy_stream.cuda_stream, ring_events=8 * per_packet)
FL = EventStream.PUSH_BORROW | EventStream.PUSH_SORTED
L = _capi.lib(); FE, ST = fe._h, st._s
om = (C.c_double * 3)(0.3, -0.2, 0.5); rc_ = (C.c_double * 1)(); rg = (C.c_double * 3)()
k_ready = C.c_int(0); p_ev, n_ev, t_pk, f_long = C.c_void_p(), C.c_size_t(0), _capi.Stamp(), C.c_int(0)
mk = lambda: torch.cuda.Event(enable_timing=True)
rows = []; out = 0; slot = 0; msg = 0
e_base = mk(); 
def step(rec):
    global msg, out, slot
    lo, hi = int(edges[msg]), int(edges[msg + 1]); msg += 1
    h0 = time.perf_counter()
    L.cmaxb_stream_push_ex(ST, C.c_void_p(base + 16 * lo), hi - lo, FL, C.byref(k_ready))
    ec = mk(); ec.record(copy_stream)
    got = 0
    while True:
        rc = L.cmaxb_stream_next_packet_device(ST, C.byref(p_ev), C.byref(n_ev), C.byref(t_pk), C.byref(f_long))
        if rc == 1: break
        L.cmaxb_fe_select_packet(FE, slot % 4); slot += 1
        L.cmaxb_stream_wait_copied(ST, C.c_void_p(stream.cuda_stream))
        ew = mk(); ew.record(stream)
        L.cmaxb_fe_set_packet_view(FE, p_ev, n_ev.value, float(t_pk.sec) + 1e-9 * float(t_pk.nsec))
        ep = mk(); ep.record(stream)
        L.cmaxb_fe_eval_launch(FE, om, 1, 1)
        ee = mk(); ee.record(stream)
        h1 = time.perf_counter()
        out += 1; got += 1
        if out >= 3:
            L.cmaxb_fe_eval_fetch(FE, rc_, rg); out -= 1
        if rec: rows.append((ec, ew, ep, ee, h0, h1, time.perf_counter(), hi - lo))
    return got
n = 0
while n < 6: n += step(False)
while out: L.cmaxb_fe_eval_fetch(FE, rc_, rg); out -= 1
torch.cuda.synchronize()
e_base.record(stream); copy_stream.wait_stream(stream)
hb = time.perf_counter()
n = 0
while n < steps: n += step(True)
while out: L.cmaxb_fe_eval_fetch(FE, rc_, rg); out -= 1
torch.cuda.synchronize()
print("step  events  host_issue  host_launched host_after_fetch | copy_done  wait_passed  prep_done  eval_done   (us since base)")
for i, (ec, ew, ep, ee, h0, h1, h2, m) in enumerate(rows):
    print("%3d %8d %10.1f %12.1f %14.1f | %9.1f %11.1f %10.1f %10.1f" % (i, m, (h0 - hb) * 1e6, (h1 - hb) * 1e6, (h2 - hb) * 1e6,
          e_base.elapsed_time(ec) * 1e3, e_base.elapsed_time(ew) * 1e3, e_base.elapsed_time(ep) * 1e3, e_base.elapsed_time(ee) * 1e3))
