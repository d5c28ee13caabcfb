# scratch GPU probe: FE parity + timing (not part of the test-suite)
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth, _capi
from cmax_slam_b200.frontend import AngVelEstimatorCMax, GRAD_ADJOINT, GRAD_DENSE
from oracle import oracle_py as O

def run(name, scale):
    pk = synth.fe_config(name, scale)
    a = O.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    w = pk.omega_true + np.array([0.2, -0.1, 0.15])
    t = time.time(); ro = O.fe_eval(a, w, True, images=True, cells=True); t_cpu = time.time() - t
    for mode in (GRAD_DENSE, GRAD_ADJOINT):
        fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, grad_mode=mode, max_hypotheses=4)
        fe.set_packet(pk.events, pk.t_ref_sec)
        cells = fe.warped_cells(w)
        c, g = fe.eval(w, True)
        c0, _ = fe.eval(w, False)
        iwe_raw = fe.computeImageOfWarpedEvents(w, blurred=False)
        iwe, d = fe.computeImageOfWarpedEvents(w, with_deriv=True, blurred=True)
        print(name, "mode", mode, "n", len(pk.events), "cells mismatch", int((cells != ro["cells"]).sum()),
              "| C gpu", c, "C0", c0, "oracle", ro["contrast"], "rel", abs(c - ro["contrast"]) / ro["contrast"])
        print("   grad gpu", g, "oracle", ro["grad"], "relerr", np.abs(g - ro["grad"]).max() / np.abs(ro["grad"]).max())
        print("   iwe_raw maxdiff", np.abs(iwe_raw - ro["iwe_raw"]).max(), "sum", iwe_raw.sum(dtype=np.float64), ro["n_inbounds"],
              "iwe blur maxdiff", np.abs(iwe - ro["iwe"]).max(), "deriv maxdiff", np.abs(d - ro["deriv"]).max(), "max", np.abs(ro["deriv"]).max())
        # timing
        for want in (False, True):
            for _ in range(5): fe.eval(w, want)
            t = time.time(); N = 50
            for _ in range(N): fe.eval(w, want)
            dt = (time.time() - t) / N
            print(f"   eval want_grad={want}: {dt*1e6:.1f} us  -> {len(pk.events)/dt:.3e} ev/s   (cpu oracle {t_cpu*1e3:.1f} ms)")
        fe.profile(True)
        for _ in range(20): fe.eval(w, True)
        for _ in range(20): fe.eval(w, False)
        print("   kernels:", {k: (round(v[0] / v[1] * 1e3, 2), v[1]) for k, v in fe.kernel_times().items()}, "us avg")
        fe.profile(False)
        # batch
        oms = synth.fe_hypotheses(pk, 4)
        cb, gb = fe.eval_batch(oms, True)
        cs = [fe.eval(o, True) for o in oms]
        print("   batch vs single", np.abs(cb - np.array([x[0] for x in cs])).max(), np.abs(gb - np.array([x[1] for x in cs])).max())
        fe.close()

run("C1", 0.2)
run("C1", 1.0)
run("C2", 1.0)
print("launches", _capi.launch_count())
