import sys; sys.path.insert(0,".")
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
from oracle import oracle_py as O
from oracle.gsl_fr import minimize_fr
pk = synth.fe_config("C1", scale=0.3)
a = O.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
f = lambda x: -O.fe_eval(a, x, False)["contrast"]
def fdf(x):
    r = O.fe_eval(a, x, True); return -r["contrast"], -r["grad"]
x0 = np.array([0.3, -0.5, 1.0])
x_ref, st_ref = minimize_fr(f, fdf, x0)
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut); fe.set_packet(pk.events, pk.t_ref_sec)
x, st = fe.setupProblemAndOptimize(x0)
print("lib ", x, st)
print("ref ", x_ref, {k:v for k,v in st_ref.items() if k!="trace"})
# python loop over the GPU cost
fg = lambda x: -fe.eval(x, False)[0]
def fdfg(x):
    c, g = fe.eval(x, True); return -c, -g
x2, st2 = minimize_fr(fg, fdfg, x0)
print("py+gpu", x2, {k:v for k,v in st2.items() if k!="trace"})
