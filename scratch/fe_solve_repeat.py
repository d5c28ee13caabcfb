# Run-to-run variation of the front-end solve on the device: the same packet solved 30 times (and the cost evaluated 200 times at one point)
import sys; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C1", scale=0.3)
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut)
fe.set_packet(pk.events, pk.t_ref_sec)
x0 = np.array([0.3, -0.5, 1.0])
om = np.array([0.45, -0.8, 1.7])
cs = np.array([fe.eval(om, True)[0] for _ in range(200)]); gs = np.array([fe.eval(om, True)[1] for _ in range(200)])
print("cost at a fixed point, 200 evaluations: rel spread %.2e; gradient rel spread %.2e" % ((cs.max() - cs.min()) / abs(cs.mean()), (gs.max(0) - gs.min(0)).max() / np.abs(gs).max()))
res = []
for i in range(30):
    if i % 10 == 0: fe.set_packet(pk.events, pk.t_ref_sec)      # re-binned (arbitrary order inside a tile)
    x, st = fe.setupProblemAndOptimize(x0)
    res.append((st["iterations"], st["f_evals"], st["cost_final"], x))
its = [r[0] for r in res]
print("iterations:", its)
print("final omega spread:", np.ptp(np.array([r[3] for r in res]), axis=0), " omega_true", pk.omega_true)
print("final cost min / max:", min(r[2] for r in res), max(r[2] for r in res))
for r in res:
    if r[0] != max(set(its), key=its.count): print("  odd run:", r[0], r[1], r[2], r[3])
