import sys; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C2")
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, grad_mode=1)
fe.set_packet(pk.events, pk.t_ref_sec)
w = pk.omega_true + np.array([0.2,-0.1,0.15])
for i in range(16):
    fe.eval(w, True)
fe.close()
