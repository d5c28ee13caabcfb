# Measures every BASELINE.json config on one GPU (C3/C5 per-GPU shares), prints one JSON line per config.
import sys, time, json; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
from cmax_slam_b200.backend import EventWarperCMax

def timeit(fn, n=20, warm=3):
    for _ in range(warm): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n

out = []
for name, k in (("C1", 1), ("C2", 1), ("C3", 32)):
    pk = synth.fe_config(name)
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, max_hypotheses=k)
    fe.set_packet(pk.events, pk.t_ref_sec)
    oms = synth.fe_hypotheses(pk, k, sigma=0.05)
    tg = timeit(lambda: fe.eval_batch(oms, True)); tv = timeit(lambda: fe.eval_batch(oms, False))
    out.append({"config": name, "events": len(pk.events), "image": [pk.width, pk.height], "hypotheses_per_gpu": k,
                "f+g_us": tg * 1e6, "f+g_warped_ev_s": k * len(pk.events) / tg, "value_us": tv * 1e6, "value_warped_ev_s": k * len(pk.events) / tv})
    print(json.dumps(out[-1]), flush=True)
    fe.close()

def be_case(name, n_events, K, pw, ph, seed, tile=1):
    base = synth.make_be_window(n_events // tile, K, pw, ph, seed, order=2, n_landmarks=50000)
    ev = base.events
    if tile > 1:   # cheap scale-up: repeat the event stream with jittered pixels is unnecessary for timing -- replicate in time order
        ev = np.repeat(ev, tile)
    rng = np.random.default_rng(seed)
    IGp = np.abs(rng.normal(0, 0.3, (ph, pw))).astype(np.float32)
    be = EventWarperCMax(base.sensor_width, base.sensor_height, base.lut, pw, ph, spline_order=2)
    be.set_window(ev, base.knots_xyzw, base.t0_ns, base.dt_ns, base.n_fixed, base.tnext, IGp, 0.5)
    x = rng.normal(0, 0.01, 3 * (K - base.n_fixed))
    tg = timeit(lambda: be.eval(x, True), n=5, warm=2); tv = timeit(lambda: be.eval(x, False), n=5, warm=2)
    c, g = be.eval(x, True)
    ilo, iln = be.local_iwe(x)
    cells = be.warped_cells(x)
    n_in = int((cells >= 0).sum())
    checksum = float(ilo.astype(np.float64).sum() + iln.astype(np.float64).sum())
    out.append({"config": name, "events": len(ev), "knots": K, "pano": [pw, ph], "f+g_us": tg * 1e6, "f+g_ev_s": len(ev) / tg,
                "value_us": tv * 1e6, "value_ev_s": len(ev) / tv, "contrast": c, "grad_finite": bool(np.all(np.isfinite(g))),
                "votes_sum": checksum, "inbounds": n_in, "votes_sum_rel_err": abs(checksum - n_in) / max(n_in, 1)})
    print(json.dumps(out[-1]), flush=True)
    be.close()

be_case("C4", 10_000_000, 64, 1280, 720, 4)
be_case("C5 (single GPU, all 5e7 events)", 50_000_000, 256, 4096, 2048, 5, tile=5)
