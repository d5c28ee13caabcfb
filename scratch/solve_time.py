# wall time of a whole front-end packet solve on the device (C2: 1M events), plain and with fused trials
import sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C2")
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut)
fe.set_packet(pk.events, pk.t_ref_sec)
for x0 in (pk.omega_true + np.array([0.3, -0.3, 0.5]), pk.omega_true * 0.6):
    for name, prm in (("plain", None), ("fused trials", (0.1, 0.05, 50, 1e-3, 1e-4, 1))):
        fe.setupProblemAndOptimize(x0, params=prm)
        t = time.perf_counter()
        for _ in range(5): x, st = fe.setupProblemAndOptimize(x0, params=prm)
        ms = (time.perf_counter() - t) / 5 * 1e3
        print("x0", np.round(x0, 2), "%-13s %.2f ms  iterations %d f_evals %d g_evals %d cost launches %d  |omega - true| %.1e  us/launch %.1f" % (
            name, ms, st["iterations"], st["f_evals"], st["g_evals"], st["cost_launches"], np.abs(x - pk.omega_true).max(), ms * 1e3 / st["cost_launches"]))
