import sys, time; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C2")
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, grad_mode=1, max_hypotheses=32)
fe.set_packet(pk.events, pk.t_ref_sec)
for k in (1, 2, 4, 8, 16, 32):
    oms = synth.fe_hypotheses(pk, k, sigma=0.05)
    for want in (False, True):
        for _ in range(3): fe.eval_batch(oms, want)
        t = time.perf_counter(); N = 20
        for _ in range(N): fe.eval_batch(oms, want)
        dt = (time.perf_counter() - t) / N
        print(f"k={k:2d} want_grad={want}: {dt*1e6:8.1f} us/batch  {dt*1e6/k:7.1f} us/hyp  {k*len(pk.events)/dt:.3e} warped-ev/s")
fe.profile(True)
oms = synth.fe_hypotheses(pk, 32, sigma=0.05)
fe.eval_batch(oms, True); print("phases k=32 f+g", np.round(fe.phase_times(),1))
fe.eval_batch(oms, False); print("phases k=32 value", np.round(fe.phase_times(),1))
fe.close()
