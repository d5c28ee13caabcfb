# wall time of the whole data path on the device for the synthetic sequence of tests/test_end_to_end.py
import sys, time; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.pipeline import CMaxSLAM
K_T = (120.0, 122.0, 63.0, 47.0)
w = synth.make_be_window(300000, 19, 256, 128, 31, order=2, sensor=(128, 96), K4=K_T, n_landmarks=800, knot_sigma=0.1)
for rep in range(2):
    slam = CMaxSLAM(128, 96, K_T, w.lut, num_events_per_packet=6000, dt_ang_vel=0.01, backend_time_window_size=0.2,
                    backend_sliding_window_stride=0.1, dt_knots=0.05, spline_degree=1, pano_height=128, max_update_times=30)
    slam.ang_vel = np.array([0.05, -0.05, 0.05])
    t = time.perf_counter()
    for i in range(0, len(w.events), 5000):
        slam.eventsCallback(w.events[i:i + 5000])
    dt = time.perf_counter() - t
    fe_ev = sum(s["f_evals"] + s["g_evals"] for _, _, s in slam.ang_vels if s)
    be_ev = sum(r["opt"]["f_evals"] + r["opt"]["g_evals"] for r in slam.windows)
    print(f"run {rep}: 0.9 s of events ({len(w.events)}) processed in {dt*1e3:.1f} ms: {len(slam.ang_vels)} packets ({fe_ev} FE cost evaluations), "
          f"{len(slam.windows)} windows ({be_ev} BE cost evaluations)")
    slam.close()
