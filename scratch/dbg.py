import sys; sys.path.insert(0,'.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
K_T = (60.0, 61.0, 31.5, 23.5)
w = synth.make_be_window(20000, 8, 128, 64, 9, order=2, sensor=(64,48), K4=K_T, n_landmarks=300, n_fixed=1)
for mode in (0,1):
    be = EventWarperCMax(64,48,w.lut,128,64,spline_order=2,grad_mode=mode)
    be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, None, 0.5)
    try:
        print(mode, be.eval(None, True))
    except Exception as e:
        print(mode, "ERR", e)
    be.close()
import ctypes as C
be = EventWarperCMax(64,48,w.lut,128,64,spline_order=2,grad_mode=0)
be.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, None, 0.5)
offs=[0xf8,0x100,0x128,0x130,0x138,0x148,0x150,0x158,0x160,0x1b0,0x1d0,0x1d8,0x1e0,0x1e8,0x208,0x210,0x1f0,0x1f8,0x220,0x228,0x230,0x238,0x240,0x248,0x250,0x258,0x268,0x280,0x290]
def dump(tag):
    base = be._h.value
    vals = [C.c_uint64.from_address(base+o).value for o in offs]
    print(tag, [hex(v) for v in vals])
dump("after set_window")
be.eval(None, False); dump("after eval f")
be.eval(None, True); dump("after eval g")
