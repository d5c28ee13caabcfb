# How far do two runs of the pipeline drift apart when the cost differs at the 1e-9 level (as device vs oracle do)?
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from oracle import oracle_py as O, pgo_py
from test_pgo import _scenario, _cut, _qdist
import oracle.gsl_fr as G
order = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w, stamps, ws, ev_t = _scenario(order)
def run(noise_seed):
    rng = np.random.default_rng(noise_seed) if noise_seed is not None else None
    orig = O.be_eval
    def noisy(a, x=None, want_grad=True, images=False, cells=False):
        r = orig(a, x, want_grad, images, cells)
        if rng is not None:
            r["contrast"] *= 1 + rng.normal(0, 1e-9)
            if r["grad"] is not None: r["grad"] = r["grad"] * (1 + rng.normal(0, 1e-7, len(r["grad"])))
        return r
    pgo_py.O.be_eval = noisy
    ref = pgo_py.PipelineOracle(w.lut, 64, 48, 128, 64, order, 0.05, 0.2, 0.1, max_update_times=20, min_num_ev=100)
    for s, v in zip(stamps, ws): ref.push(s, v)
    reps = []
    for win in range(3):
        ev = _cut(w, ev_t, ref.t_win_beg, ref.t_win_end)
        reps.append((ref.process(ev), ref.knots.copy(), ref.IG.sum(), ref.times.astype(np.int64).sum()))
    pgo_py.O.be_eval = orig
    return reps, ref
base, ref0 = run(None)
for seed in range(6):
    other, ref1 = run(seed)
    for win in range(3):
        (ra, ka, iga, ta), (rb, kb, igb, tb) = base[win], other[win]
        n_seg = len(ka) - order + 1
        worst = 0
        for u in np.linspace(0.05, n_seg - 1.05, 25):
            t_ns = ref0.traj_t_beg_ns + int(u * ref0.traj_dt_ns)
            a = O.spline_eval(order, ka, ref0.traj_t_beg_ns, ref0.traj_dt_ns, t_ns, want_J=False)[0]
            b = O.spline_eval(order, kb, ref0.traj_t_beg_ns, ref0.traj_dt_ns, t_ns, want_J=False)[0]
            worst = max(worst, _qdist(a, b))
        print(seed, win, "curve %.2e ctrl %.2e cost %.2e alpha %.2e latest %.2e IG %.2e times %.2e iters %d/%d" % (
            worst, _qdist(ka, kb), abs(ra["opt"]["cost_final"] - rb["opt"]["cost_final"]) / abs(ra["opt"]["cost_final"]),
            abs(ra["alpha"] - rb["alpha"]) / max(abs(ra["alpha"]), 1e-12), _qdist(ra["pose_latest"][1], rb["pose_latest"][1]),
            abs(iga - igb) / iga, abs(ta - tb) / max(ta, 1), ra["opt"]["iterations"], rb["opt"]["iterations"]))
