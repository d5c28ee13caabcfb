# %globaltimer phase stamps of the fused front-end kernel (synchronous evaluations, whole-GPU grid) + the results, so that
# build variants can be compared for speed AND value.
import sys; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
np.set_printoptions(linewidth=200)
for name, scale in (("C1", 1.0), ("C2", 1.0)):
    pk = synth.fe_config(name, scale)
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, grad_mode=1)
    fe.set_packet(pk.events, pk.t_ref_sec)
    w = pk.omega_true + np.array([0.2,-0.1,0.15])
    for i in range(5): c, g = fe.eval(w, True)
    print(name, "contrast %.12f grad" % c, g, fe.launch_info())
    fe.profile(True)
    for want in (True, False):
        for i in range(3):
            fe.eval(w, want)
            print(name, want, np.round(fe.phase_times(), 2))
    fe.close()
