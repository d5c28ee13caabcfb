"""The C++ reference-side binding of the widened rows (include/cmax_b200_pipeline.hpp) compiles with -Wall -Werror,
links against libcmax_b200.so and -- driven from C++ alone, no Python binding in the loop -- cuts the same packets and
produces the same control poses as the Python mirror (host-only path: no device needed)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cmax_slam_b200 import build, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SRC = r'''
#include "cmax_b200_pipeline.hpp"
using namespace cmaxb_pipeline;
// feeds `n` events in messages of `msg`, counts packets; angular velocity = (1, -2, 0.5) * packet index parity;
// returns the control poses of the bookkeeping-only window solver
extern "C" int drive(const cmaxb_event* ev, size_t n, size_t msg, double dt_ang_vel, int per_packet, int order,
                     int* n_packets, long long* n_packet_events, int* n_windows, double* ctrl_xyzw, int cap, int* n_ctrl) {
  try {
    EventStore store(dt_ang_vel, per_packet, 1);
    cmaxb_pgo_cfg cfg{};
    cfg.spline_order = order; cfg.dt_knots = 0.05; cfg.time_window_size = 0.2; cfg.sliding_window_stride = 0.1;
    cfg.y_angle_deg = 0.0; cfg.max_update_times = 10; cfg.min_num_ev_per_win = 1e18; cfg.use_opt_params = 0;
    WindowSolver solver(cfg, nullptr);
    *n_packets = 0; *n_packet_events = 0; *n_windows = 0;
    for (size_t i = 0; i < n; i += msg) {
      store.push(ev + i, (n - i < msg) ? n - i : msg);
      Packet pk;
      while (store.next_packet(&pk)) {
        const double s = (*n_packets % 2) ? 1.0 : -1.0;
        const double w[3] = {1.0 * s, -2.0 * s, 0.5};
        solver.push_ang_vel(pk.time_packet, w);
        *n_packets += 1;
        *n_packet_events += (long long)pk.events.size;
        *n_windows += solver.run_ready_windows(store, [](const cmaxb_pgo_report&) {});
      }
    }
    // a window the store does not reach yet is "not ready" (false), not an error
    Span far;
    cmaxb_stamp tb{ev[n - 1].sec + 10, ev[n - 1].nsec}, te{ev[n - 1].sec + 11, ev[n - 1].nsec};
    if (store.window_events(tb, te, &far)) return -101;
    std::vector<double> q = solver.ctrl_poses_xyzw();
    *n_ctrl = (int)(q.size() / 4);
    if ((int)q.size() > 4 * cap) return -100;
    for (size_t i = 0; i < q.size(); ++i) ctrl_xyzw[i] = q[i];
    return 0;
  } catch (const Error& e) {
    return e.code;
  }
}
'''


@pytest.fixture(scope="module")
def drv(tmp_path_factory):
    d = tmp_path_factory.mktemp("pipe")
    src = d / "drive.cpp"
    src.write_text(_SRC)
    out = d / "libdrive.so"
    lib = build.build()
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", str(out), str(src), lib, f"-Wl,-rpath,{os.path.dirname(lib)}"])
    return C.CDLL(str(out))


@pytest.mark.parametrize("order", [2, 4])
def test_cpp_driver_matches_python_mirror(drv, order):
    from cmax_slam_b200.backend import PoseGraphOptimizerCMax
    from cmax_slam_b200.stream import EventStream
    from cmax_slam_b200._capi import CmaxbError
    rng = np.random.default_rng(12)
    n = 120000
    t_ns = (synth.EPOCH_SEC * 1_000_000_000 + 5_000 + np.cumsum(rng.exponential(1e9 / 1.5e5, n))).astype(np.int64)
    ev = np.zeros(n, synth.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 64, n); ev["y"] = rng.integers(0, 48, n)
    ev["sec"] = t_ns // 1_000_000_000; ev["nsec"] = t_ns % 1_000_000_000
    npk, nev, nwin, nctrl = C.c_int(0), C.c_longlong(0), C.c_int(0), C.c_int(0)
    ctrl = np.zeros((64, 4))
    drv.drive.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                          C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int)]
    rc = drv.drive(ev.ctypes.data, n, 3000, 0.01, 1500, order, C.byref(npk), C.byref(nev), C.byref(nwin),
                   ctrl.ctypes.data_as(C.POINTER(C.c_double)), 64, C.byref(nctrl))
    assert rc == 0
    # the same scenario through the Python mirror
    s = EventStream(0.01, 1500, 1)
    pgo = PoseGraphOptimizerCMax(None, order, 0.05, 0.2, 0.1, min_num_ev_per_win=1e18)
    k = tot = wins = 0
    for i in range(0, n, 3000):
        s.eventsCallback(ev[i:i + 3000])
        while True:
            p = s.next_packet()
            if p is None:
                break
            sg = 1.0 if k % 2 else -1.0
            pgo.pushAngVel(p[1], [1.0 * sg, -2.0 * sg, 0.5])
            k += 1; tot += len(p[0])
            while True:
                tb, te, ready = pgo.window()
                if not ready:
                    break
                try:
                    w = s.window_events(tb, te)
                except CmaxbError:
                    break
                pgo.processTimeWindow(w)
                wins += 1
    q, _, _ = pgo.ctrl_poses()
    assert (npk.value, nev.value, nwin.value, nctrl.value) == (k, tot, wins, len(q))
    assert k > 50 and wins >= 4 and len(q) >= 8
    assert np.array_equal(ctrl[: len(q)], q)                  # same library, same inputs: bit-identical
