import sys; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
for name, scale in (("C1", 0.2), ("C2", 1.0)):
    pk = synth.fe_config(name, scale)
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, grad_mode=1)
    fe.set_packet(pk.events, pk.t_ref_sec)
    w = pk.omega_true + np.array([0.2,-0.1,0.15])
    for i in range(4):
        fe.eval(w, True)
    for i in range(4):
        fe.eval(w, False)
    fe.close()
