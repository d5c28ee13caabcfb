"""The library's event ingestion (csrc/stream.cu), trajectory initialisation (csrc/traj_init.cu) and window bookkeeping
(csrc/pgo.cu) against the reference's OWN node classes -- src/frontend/ang_vel_estimator.cpp and
src/backend/pose_graph_optimizer.cpp (+ trajectory.cpp, event_pano_warper.cpp, ...) compiled unmodified with the stand-in
headers of oracle/stubs/ and driven without ROS (oracle/ref_node_shim.cpp): the reference receives the events ONE BY ONE through
pushEvent with the Run() loop body after each, the library receives them in messages.  The two GSL solves are stand-ins on both
sides (front-end: next row of a table; back-end: control poses unchanged), so what is compared is exactly the host logic:
packets (stamp, events), windows (cursors, events), control poses, index arithmetic, latest pose.  Host code: no device needed;
skipped when oracle/_ref was not built (reference tree absent)."""
import numpy as np
import pytest

from cmax_slam_b200 import synth


def _stream(n, seed, rate_hz=1.5e5, gap_at=None, gap_s=0.0):
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0 / rate_hz, n)
    if gap_at is not None:
        dt[gap_at] += gap_s
    t_ns = (synth.EPOCH_SEC * 1_000_000_000 + 7_654_321 + np.cumsum(dt * 1e9)).astype(np.int64)
    ev = np.zeros(n, synth.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 64, n); ev["y"] = rng.integers(0, 48, n)
    ev["sec"] = t_ns // 1_000_000_000; ev["nsec"] = t_ns % 1_000_000_000
    return ev


def _qdist(a, b):
    a, b = np.atleast_2d(a), np.atleast_2d(b)
    c = np.minimum(np.linalg.norm(a - b, axis=1), np.linalg.norm(a + b, axis=1))
    return float((4 * np.arcsin(np.clip(c / 2, 0, 1))).max())


@pytest.mark.parametrize("degree,fe_rate,per_packet,msg,gap", [(1, 1, 2000, 3000, None), (3, 1, 1500, 977, None), (1, 2, 1200, 4096, None),
                                                             (1, 1, 2000, 2500, (60000, 0.25))])
def test_library_host_logic_matches_reference_node_classes(oracle, degree, fe_rate, per_packet, msg, gap):
    if not oracle.have_ref_node():
        pytest.skip("oracle/_ref/libref_node.so not built (reference tree absent)")
    from cmax_slam_b200._capi import CmaxbError
    from cmax_slam_b200.backend import PoseGraphOptimizerCMax
    from cmax_slam_b200.stream import EventStream
    K4 = (60.0, 61.0, 31.5, 23.5)
    ev = _stream(130000, 21 + degree, gap_at=None if gap is None else gap[0], gap_s=0.0 if gap is None else gap[1])
    rng = np.random.default_rng(4)
    omegas = rng.normal(0, 1.5, (37, 3))
    order = 4 if degree == 3 else 2
    ref = oracle.RefNode(64, 48, K4, omegas, dt_ang_vel=0.01, num_events_per_packet=per_packet, fe_sample_rate=fe_rate, dt_knots=0.05,
                         spline_degree=degree, y_angle=15.0)
    ref.events(ev)                                             # one pushEvent per (subsampled) event
    n_pk, n_win, n_stored = ref.counts()
    assert n_pk > 40 and n_win >= 3

    s = EventStream(0.01, per_packet, fe_rate)
    pgo = PoseGraphOptimizerCMax(None, order, 0.05, 0.2, 0.1, y_angle_deg=15.0)
    packets, windows = [], []
    k_solved = 0
    for i in range(0, len(ev), msg):
        s.eventsCallback(ev[i:i + msg])
        while True:
            p = s.next_packet()
            if p is None:
                break
            e, tp, too_long = p
            if too_long:
                om = np.zeros(3)
                packets.append((tp, -1, None))
            else:
                om = omegas[k_solved % len(omegas)]
                k_solved += 1
                packets.append((tp, len(e), e.copy()))
            pgo.pushAngVel(tp, om)
            while True:
                tb, te, ready = pgo.window()
                if not ready:
                    break
                try:
                    wev = s.window_events(tb, te)
                except CmaxbError:
                    break
                rep = pgo.processTimeWindow(wev)
                q, _, _ = pgo.ctrl_poses()
                windows.append((tb, te, wev.copy(), rep, q))
    # the reference saw the whole stream event by event; the library may hold a few trailing windows back only if its message
    # boundaries delayed them -- at the end of the stream both have processed the same number
    assert len(packets) == n_pk and len(windows) == n_win, (len(packets), n_pk, len(windows), n_win)
    for i, (tp, n, e) in enumerate(packets):
        v, h = ref.packet(i)
        assert (v[0], v[1]) == tuple(tp) and v[2] == n, (i, v, tp, n)
        if n > 0:
            assert (v[3], v[4], v[5], v[6]) == (int(e["sec"][0]), int(e["nsec"][0]), int(e["sec"][-1]), int(e["nsec"][-1]))
            assert h == oracle.hash_events(e), i
    for i, (tb, te, wev, rep, q) in enumerate(windows):
        v, h, latest_q, knots = ref.window(i)
        assert (v[0], v[1], v[2], v[3]) == (*tb, *te) and v[4] == len(wev), (i, v[:5], tb, te, len(wev))
        assert h == oracle.hash_events(wev), i
        assert (v[9], v[10], v[11], v[12]) == (rep["n_ctrl_poses"], rep["idx_cp_traj_beg"], rep["idx_cp_opt_beg"], rep["num_cp_opt"]), (i, v[9:13], rep)
        assert (v[13], v[14]) == tuple(rep["pose_latest"][0])
        assert _qdist(rep["pose_latest"][1], latest_q) <= 1e-9, i
        assert q.shape == knots.shape and _qdist(q, knots) <= 1e-8, (i, _qdist(q, knots))
    assert s.state()["n_stored"] == n_stored
    if gap is not None:
        assert any(n == -1 for _, n, _ in packets)            # the silent gap produced a packet the reference does not solve
    ref.close()
    pgo.close()
    s.close()


@pytest.mark.parametrize("tag,degree", [("lin", 1), ("cub", 3)])
def test_library_host_logic_matches_golden_of_reference_node_classes(oracle, golden, tag, degree):
    """The same comparison against the committed outputs of the reference's node classes (tests/golden/node_firstparty.npz,
    generated by tests/golden/make_golden.py through oracle/_ref/libref_node.so): runs without the reference tree."""
    from cmax_slam_b200._capi import CmaxbError
    from cmax_slam_b200.backend import PoseGraphOptimizerCMax
    from cmax_slam_b200.stream import EventStream
    g = golden("node_firstparty.npz")
    ev = _stream(130000, 21 + degree)
    omegas = np.random.default_rng(4).normal(0, 1.5, (37, 3))
    s = EventStream(0.01, 2000, 1)
    pgo = PoseGraphOptimizerCMax(None, 4 if degree == 3 else 2, 0.05, 0.2, 0.1, y_angle_deg=15.0)
    pk_rows, pk_hash, win_rows, win_hash, latest, knots = [], [], [], [], [], []
    k = 0
    for i in range(0, len(ev), 3333):
        s.eventsCallback(ev[i:i + 3333])
        while True:
            p = s.next_packet()
            if p is None:
                break
            e, tp, too_long = p
            assert not too_long
            pk_rows.append([tp[0], tp[1], len(e), int(e["sec"][0]), int(e["nsec"][0]), int(e["sec"][-1]), int(e["nsec"][-1])])
            pk_hash.append(oracle.hash_events(e))
            pgo.pushAngVel(tp, omegas[k % len(omegas)])
            k += 1
            while True:
                tb, te, ready = pgo.window()
                if not ready:
                    break
                try:
                    wev = s.window_events(tb, te)
                except CmaxbError:
                    break
                rep = pgo.processTimeWindow(wev)
                win_rows.append([*tb, *te, len(wev), int(wev["sec"][0]), int(wev["nsec"][0]), int(wev["sec"][-1]), int(wev["nsec"][-1]),
                                 rep["n_ctrl_poses"], rep["idx_cp_traj_beg"], rep["idx_cp_opt_beg"], rep["num_cp_opt"], *rep["pose_latest"][0]])
                win_hash.append(oracle.hash_events(wev))
                latest.append(rep["pose_latest"][1])
                knots.append(pgo.ctrl_poses()[0])
    assert np.array_equal(np.array(pk_rows, dtype=np.int64), g[f"{tag}_packets"])
    assert np.array_equal(np.array(pk_hash, dtype=np.uint64), g[f"{tag}_packet_hash"])
    assert np.array_equal(np.array(win_rows, dtype=np.int64), g[f"{tag}_windows"])
    assert np.array_equal(np.array(win_hash, dtype=np.uint64), g[f"{tag}_window_hash"])
    assert _qdist(np.array(latest), g[f"{tag}_latest_q"]) <= 1e-9
    for i, q in enumerate(knots):
        assert q.shape == g[f"{tag}_knots{i}"].shape and _qdist(q, g[f"{tag}_knots{i}"]) <= 1e-8, i
    assert s.state()["n_stored"] == int(g[f"{tag}_n_stored"])
    pgo.close()
    s.close()
