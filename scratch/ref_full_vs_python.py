# The complete reference pipeline on the CPU (real translation units, GSL stand-in) vs the Python restatement pipeline
import sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.stream import EventStream
from oracle import oracle_py as O
from oracle.gsl_fr import minimize_fr
from oracle.pgo_py import PipelineOracle
K_T = (120.0, 122.0, 63.0, 47.0)
n_ev = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
w = synth.make_be_window(n_ev, 9, 256, 128, 31, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)
t = time.time()
ref = O.RefNode(128, 96, K_T, np.zeros((1, 3)), dt_ang_vel=0.01, num_events_per_packet=6000, dt_knots=0.05, spline_degree=1, pano_height=128,
                min_ev_rate=10, max_update_times=30, full=True)
ref.events(w.events)
n_pk, n_win, _ = ref.counts()
print("reference (real TUs):", n_pk, "packets,", n_win, "windows in %.1f s" % (time.time() - t))
# Python restatement pipeline
t = time.time()
s = EventStream(0.01, 6000, 1)
pgo = PipelineOracle(w.lut, 128, 96, 256, 128, 2, 0.05, 0.2, 0.1, max_update_times=30, min_num_ev=int(0.2 * 10 / 1))
om = np.zeros(3); avs = []; wins = []
for i in range(0, len(w.events), 5000):
    s.eventsCallback(w.events[i:i + 5000])
    while True:
        pk = s.next_packet()
        if pk is None: break
        ev, tp, tl = pk
        a = O.fe_args(ev.copy(), tp[0] + 1e-9 * tp[1], w.lut, 128, 96, K_T)
        f = lambda x: -O.fe_eval(a, x, False)["contrast"]
        def fdf(x):
            r = O.fe_eval(a, x, True); return -r["contrast"], -r["grad"]
        om, st = minimize_fr(f, fdf, om)
        avs.append((tp, om.copy()))
        pgo.push(tp, om)
        while pgo.init and sorted(pgo.av) and sorted(pgo.av)[-1] > pgo.t_win_end:
            try:
                evw = s.window_events(pgo.t_win_beg, pgo.t_win_end)
            except Exception:
                break
            wins.append((pgo.process(evw.copy()), pgo.knots.copy()))
print("python restatement:", len(avs), "packets,", len(wins), "windows in %.1f s" % (time.time() - t))
worst = 0
for i, (tp, o) in enumerate(avs[:n_pk]):
    v, _ = ref.packet(i)
    wr = ref.packet_omega(i)
    assert (v[0], v[1]) == tuple(tp), (i, v, tp)
    worst = max(worst, np.abs(wr - o).max())
    if i < 5 or np.abs(wr - o).max() > 1e-9: print(i, wr, o, np.abs(wr - o).max())
print("max |omega_ref - omega_py| over packets:", worst)
for i, (rep, kn) in enumerate(wins[:n_win]):
    v, h, lq, knots = ref.window(i)
    d = np.minimum(np.linalg.norm(knots - kn, axis=1), np.linalg.norm(knots + kn, axis=1)).max()
    print("window", i, "idx", v[9:13], (rep["n_ctrl_poses"], rep["idx_cp_traj_beg"], rep["idx_cp_opt_beg"], rep["num_cp_opt"]), "ctrl pose diff", d)
IG, times = ref.get_map(256, 128)
print("IG sum ref %.3f py %.3f ; max diff %.3g ; times equal %s" % (IG.sum(), pgo.IG.sum(), np.abs(IG - pgo.IG).max(), np.array_equal(times, pgo.times)))
