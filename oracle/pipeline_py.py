"""TEST INFRASTRUCTURE.  The whole CMax-SLAM data path as a composition of the Python / oracle restatements: event ingestion
(the library's stream, pinned in tests/test_node_firstparty.py), front-end solve = oracle/gsl_fr.py over the oracle cost
(warm-started from the previous packet, zero when the packet spans more than 10 dt), back-end = oracle/pgo_py.py.  Checked
bit for bit against the reference's own translation units run as a whole (oracle/_ref/libref_full.so,
tests/test_pipeline_firstparty.py)."""
import numpy as np

from . import oracle_py as O
from .gsl_fr import minimize_fr
from .pgo_py import PipelineOracle


def run(events, lut, W, H, K4, PW, PH, order, *, dt_ang_vel=0.01, num_events_per_packet=6000, dt_knots=0.05, win_size=0.2, win_stride=0.1,
        max_update_times=30, min_ev_rate=10, msg=5000, fe_solver=None):
    """Returns dict(packets=[(stamp, omega)], windows=[(report, knots)], IG, times).  fe_solver(f, fdf, x0) -> x replaces
    oracle/gsl_fr.py (e.g. the library's cmaxb_optimize_callback)."""
    from cmax_slam_b200.stream import EventStream
    s = EventStream(dt_ang_vel, num_events_per_packet, 1)
    pgo = PipelineOracle(lut, W, H, PW, PH, order, dt_knots, win_size, win_stride, max_update_times=max_update_times,
                         min_num_ev=int(win_size * min_ev_rate / 1))
    om = np.zeros(3)
    packets, windows = [], []
    for i in range(0, len(events), msg):
        s.eventsCallback(events[i:i + msg])
        while True:
            pk = s.next_packet()
            if pk is None:
                break
            ev, tp, too_long = pk
            if too_long:
                om = np.zeros(3)
            else:
                a = O.fe_args(ev.copy(), float(tp[0]) + 1e-9 * float(tp[1]), lut, W, H, K4)
                f = lambda x, a=a: -O.fe_eval(a, x, False)["contrast"]

                def fdf(x, a=a):
                    r = O.fe_eval(a, x, True)
                    return -r["contrast"], -r["grad"]

                om = fe_solver(f, fdf, om) if fe_solver else minimize_fr(f, fdf, om)[0]
            packets.append((tp, om.copy()))
            pgo.push(tp, om)
            while pgo.init and pgo.av and max(pgo.av) > pgo.t_win_end:
                try:
                    evw = s.window_events(pgo.t_win_beg, pgo.t_win_end)
                except Exception:
                    break
                windows.append((pgo.process(evw.copy()), pgo.knots.copy()))
    s.close()
    return {"packets": packets, "windows": windows, "IG": pgo.IG.copy(), "times": pgo.times.copy()}
