"""The whole CMax-SLAM data path over the C ABI, mirroring the reference node's wiring (src/cmax_slam.cpp:13-161):

    eventsCallback -> AngVelEstimator::pushEvent -> [packet complete] processEventPacket (front-end solve)
                   -> PoseGraphOptimizer::pushAngVel -> [window covered] getEventSubset + processTimeWindow + slideWindow

Every piece is libcmax_b200.so: cmaxb_stream_* (event store, packet / window cutting), cmaxb_fe_* (packet upload +
Fletcher-Reeves solve on the GPU), cmaxb_pgo_* / cmaxb_be_* (trajectory initialisation, window solve, map upkeep on the
GPU).  ROS I/O (topics, camera_info, publishers) is the caller's; this class is what sits between them.  The reference
runs the back-end in its own thread (src/cmax_slam.cpp:92); here a window is processed as soon as both the front-end's
angular velocities and the event store cover it, which is the same data dependency without the race."""
import numpy as np

from .backend import EventWarperCMax, PoseGraphOptimizerCMax
from .frontend import AngVelEstimatorCMax
from .stream import EventStream


class CMaxSLAM:
    def __init__(self, width, height, K4, lut_xyz, *, num_events_per_packet=30000, dt_ang_vel=0.02, frontend_blur_sigma=1.0,
                 event_batch_size=100, frontend_event_sample_rate=1, contrast_measure=0, backend_time_window_size=0.2,
                 backend_sliding_window_stride=0.1, backend_blur_sigma=1.0, backend_event_sample_rate=1, dt_knots=0.1,
                 spline_degree=1, pano_height=1024, Y_angle=0.0, backend_min_ev_rate=10, max_update_times=10, device=0,
                 run_backend=True):
        order = 4 if spline_degree == 3 else 2
        self.stream = EventStream(dt_ang_vel, num_events_per_packet, frontend_event_sample_rate)
        self.fe = AngVelEstimatorCMax(width, height, K4, lut_xyz, blur_sigma=frontend_blur_sigma, event_batch_size=event_batch_size,
                                      contrast_measure=contrast_measure, device=device)
        self.be = self.pgo = None
        if run_backend:
            self.be = EventWarperCMax(width, height, lut_xyz, 2 * pano_height, pano_height, blur_sigma=backend_blur_sigma,
                                      event_batch_size=event_batch_size, event_sample_rate=backend_event_sample_rate,
                                      spline_order=order, contrast_measure=contrast_measure, device=device)
            self.be.resetIG()
            # min_num_ev_per_win_ is a size_t in the reference (pose_graph_optimizer.cpp:64-67): truncation
            min_ev = int(backend_time_window_size * backend_min_ev_rate / (backend_event_sample_rate * frontend_event_sample_rate))
            self.pgo = PoseGraphOptimizerCMax(self.be, order, dt_knots, backend_time_window_size, backend_sliding_window_stride,
                                              y_angle_deg=Y_angle, max_update_times=max_update_times, min_num_ev_per_win=min_ev)
        self.ang_vel = np.zeros(3)                 # ang_vel_: warm start of the next packet (ang_vel_estimator.h)
        self.ang_vels = []                         # (stamp, omega, solver stats) per packet
        self.windows = []                          # report per back-end window
        self._last_event_ts = None

    def close(self):
        for o in (self.pgo, self.be, self.fe, self.stream):
            if o is not None:
                o.close()
        self.pgo = self.be = self.fe = self.stream = None

    def eventsCallback(self, msg_events):
        """One event message (structured array of 16-byte dvs_msgs::Event records)."""
        if len(msg_events) == 0:
            return
        self.stream.eventsCallback(msg_events)
        self._last_event_ts = (int(msg_events["sec"][-1]), int(msg_events["nsec"][-1]))
        while True:
            pk = self.stream.next_packet()
            if pk is None:
                break
            events, t_packet, too_long = pk
            stats = None
            if too_long:                           # ang_vel_estimator.cpp:109-114
                self.ang_vel = np.zeros(3)
            else:                                  # processEventPacket: solve, warm-started from the previous estimate
                self.fe.set_packet(events, float(t_packet[0]) + 1e-9 * float(t_packet[1]))
                self.ang_vel, stats = self.fe.setupProblemAndOptimize(self.ang_vel)
            self.ang_vels.append((t_packet, self.ang_vel.copy(), stats))
            if self.pgo is not None:
                self.pgo.pushAngVel(t_packet, self.ang_vel)
                self._run_backend()

    def _run_backend(self):
        while True:
            tb, te, ready = self.pgo.window()
            if not ready:
                return
            ev = self.stream.window_events(tb, te)
            if ev is None:                        # the event store does not reach the end of the window yet
                return
            self.windows.append(self.pgo.processTimeWindow(ev))

    def trajectory(self):
        """(control poses xyzw, t0_ns, dt_ns) of the back-end spline so far."""
        return self.pgo.ctrl_poses()

    def getIG(self):
        return self.be.getIG()[0]
