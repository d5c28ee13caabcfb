# Throughput probe: two handles on two streams, each launching the fused kernel on a fraction of the GPU
# (CMAXB_FE_GRID_FRACTION), vs one handle on the whole GPU.  Wall clock over many evaluations, L2 warm.
import os, sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C2")
om = synth.fe_hypotheses(pk, 4, seed=3, sigma=0.05)
def run(n_handles, depth, n=400):
    fes = []
    for i in range(n_handles):
        fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut)
        fe.set_packet(pk.events, pk.t_ref_sec)
        fes.append(fe)
    res = []
    def loop(count):
        out = [0] * n_handles
        for i in range(count):
            h = i % n_handles
            fes[h].eval_launch(om[i % 4][None, :], True)
            out[h] += 1
            if out[h] >= depth:
                res.append(fes[h].eval_fetch()[0][0]); out[h] -= 1
        for h in range(n_handles):
            while out[h]:
                res.append(fes[h].eval_fetch()[0][0]); out[h] -= 1
    loop(40)
    t = time.perf_counter(); loop(n); dt = (time.perf_counter() - t) / n
    for fe in fes: fe.close()
    return dt * 1e6, res[-1]
frac = os.environ.get("CMAXB_FE_GRID_FRACTION", "1")
for nh, depth in ((1, 1), (1, 2), (2, 1), (2, 2), (3, 2)):
    us, c = run(nh, depth)
    print(f"fraction {frac}: {nh} handle(s), depth {depth}: {us:.1f} us per evaluation ({len(pk.events)/us*1e6:.3e} ev/s), contrast {c:.6f}")
