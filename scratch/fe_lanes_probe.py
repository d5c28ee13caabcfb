# Throughput probe of the fused front-end kernel: evaluations per second of the C2 packet (contrast + gradient) for
# several lane counts / pipeline depths (one handle; cmaxb_fe_eval_launch round-robins over the lanes).  Wall clock over
# many evaluations; "rot" = number of distinct resident packets the evaluations rotate over (working set > L2 when large).
import os, sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C2")
om = synth.fe_hypotheses(pk, 4, seed=3, sigma=0.05)
tag = os.environ.get("PROBE_TAG", "")
def run(lanes, depth, rot, want_grad=True, n=600):
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, lanes=lanes, packet_slots=rot)
    for s in range(rot):
        fe.select_packet(s)
        ev = pk.events.copy()
        fe.set_packet(ev, pk.t_ref_sec)
    res = []
    def loop(count):
        out = 0
        for i in range(count):
            fe.select_packet(i % rot)
            fe.eval_launch(om[i % 4][None, :], want_grad)
            out += 1
            if out >= depth:
                res.append(fe.eval_fetch()[0][0]); out -= 1
        while out:
            res.append(fe.eval_fetch()[0][0]); out -= 1
    loop(60)
    t = time.perf_counter(); loop(n); dt = (time.perf_counter() - t) / n
    info = fe.launch_info()
    fe.close()
    return dt * 1e6, res[-1], info
import sys as _s
CASES = ((1, 1, 1), (1, 2, 1), (2, 4, 1), (3, 6, 1), (4, 8, 1), (3, 6, 6)) if len(_s.argv) < 2 else ((3, 6, 1), (3, 6, 6))
for lanes, depth, rot in CASES:
    us, c, info = run(lanes, depth, rot)
    print(f"{tag} lanes {lanes} depth {depth} rot {rot}: {us:.1f} us/eval f+g ({len(pk.events)/us*1e6:.3e} ev/s) contrast {c:.6f} grid {info['grid_full']}/{info['grid_lane']} tma {info['tma']}", flush=True)
us, c, info = run(3, 6, 1, want_grad=False)
print(f"{tag} lanes 3 depth 6 value-only: {us:.1f} us/eval", flush=True)
us, c, info = run(1, 1, 1, want_grad=False)
print(f"{tag} lanes 1 depth 1 value-only: {us:.1f} us/eval", flush=True)
