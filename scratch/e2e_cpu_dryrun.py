# CPU dry run of tests/test_end_to_end.py with oracle pieces (threshold tuning without a GPU)
import sys; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.stream import EventStream
from oracle import oracle_py as O
from oracle.gsl_fr import minimize_fr
from oracle.pgo_py import PipelineOracle
K_T = (120.0, 122.0, 63.0, 47.0)
w = synth.make_be_window(300000, 19, 256, 128, 31, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)
s = EventStream(0.01, 6000, 1)
pgo = PipelineOracle(w.lut, 128, 96, 256, 128, 2, 0.05, 0.2, 0.1, max_update_times=30, min_num_ev=2)
om = np.array([0.05, -0.05, 0.05]); avs = []; wins = []
def backend():
    while True:
        if not pgo.init: return
        keys = sorted(pgo.av)
        if not keys or not keys[-1] > pgo.t_win_end: return
        try:
            ev = s.window_events(pgo.t_win_beg, pgo.t_win_end)
        except Exception as e:
            return
        wins.append(pgo.process(ev.copy()))
for i in range(0, len(w.events), 5000):
    s.eventsCallback(w.events[i:i+5000])
    while True:
        pk = s.next_packet()
        if pk is None: break
        ev, tp, tl = pk
        a = O.fe_args(ev.copy(), tp[0] + 1e-9 * tp[1], w.lut, 128, 96, K_T)
        f = lambda x: -O.fe_eval(a, x, False)["contrast"]
        def fdf(x):
            r = O.fe_eval(a, x, True); return -r["contrast"], -r["grad"]
        om, st = minimize_fr(f, fdf, om)
        avs.append((tp, om.copy(), st))
        pgo.push(tp, om)
        backend()
print(len(avs), len(wins))
errs = []
for (ts, o, st) in avs[2:-2]:
    t_ns = ts[0] * 10**9 + ts[1]
    seg = int((t_ns - w.t0_ns) // w.dt_ns)
    if seg < 0 or seg >= len(w.knots_xyzw) - 1: continue
    d = synth._qlog(synth._qmul(synth._qconj(w.knots_xyzw[seg][None, :]), w.knots_xyzw[seg + 1][None, :]))[0] / (w.dt_ns * 1e-9)
    frac = ((t_ns - w.t0_ns) % w.dt_ns) / w.dt_ns
    if 0.3 < frac < 0.7: errs.append(np.abs(o - d).max())
print("fe errs median", np.median(errs), "max", max(errs), len(errs))
for k, r in enumerate(wins):
    print(k, r["idx_cp_traj_beg"], r["optimized"], r.get("opt", {}).get("cost_initial"), r.get("opt", {}).get("cost_final"), r.get("opt", {}).get("iterations"))
q = pgo.knots; t0_ns, dt_ns = pgo.traj_t_beg_ns, pgo.traj_dt_ns
def rel(a, b): return synth._qmul(synth._qconj(a[None, :]), b[None, :])[0]
def truth_at(t_ns):
    seg = (t_ns - w.t0_ns) // w.dt_ns; u = ((t_ns - w.t0_ns) % w.dt_ns) / w.dt_ns
    dd = synth._qlog(rel(w.knots_xyzw[seg], w.knots_xyzw[seg + 1])[None, :])[0]
    return synth._qmul(w.knots_xyzw[seg][None, :], synth._qexp((dd * u)[None, :]))[0]
worst = 0
for i in range(2, len(q) - 2):
    t_ns = t0_ns + i * dt_ns
    if (t_ns - w.t0_ns) // w.dt_ns + 1 >= len(w.knots_xyzw): break
    ang = np.linalg.norm(synth._qlog(rel(rel(truth_at(t0_ns + 2 * dt_ns), truth_at(t_ns)), rel(q[2], q[i]))[None, :])[0])
    worst = max(worst, ang); print(i, ang)
print("worst", worst, "IG sum", pgo.IG.sum())
for (ts, o, st) in avs[10:20]:
    t_ns = ts[0] * 10**9 + ts[1]
    seg = int((t_ns - w.t0_ns) // w.dt_ns)
    d = synth._qlog(synth._qmul(synth._qconj(w.knots_xyzw[seg][None, :]), w.knots_xyzw[seg + 1][None, :]))[0] / (w.dt_ns * 1e-9)
    print(np.round(o, 3), np.round(d, 3), ((t_ns - w.t0_ns) % w.dt_ns) / w.dt_ns)
