"""Event ingestion: the library's packet / window cutter (csrc/stream.cu) against the
line-by-line per-event Python restatement of the reference (oracle/stream_py.py).  Host code: no device needed."""
import numpy as np
import pytest

from cmax_slam_b200 import synth
from cmax_slam_b200.stream import EventStream
from oracle.stream_py import StreamOracle


def _events(n, seed, rate_hz=2.0e5, gap_at=None, gap_s=0.0):
    """n events with exponential inter-arrival times (ns resolution), an optional silent gap."""
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0 / rate_hz, n)
    if gap_at is not None:
        dt[gap_at] += gap_s
    t_ns = (synth.EPOCH_SEC * 1_000_000_000 + 123_456_789 + np.cumsum(dt * 1e9)).astype(np.int64)
    ev = np.zeros(n, synth.EVENT_DTYPE)
    ev["x"] = rng.integers(0, 240, n); ev["y"] = rng.integers(0, 180, n)
    ev["sec"] = t_ns // 1_000_000_000; ev["nsec"] = t_ns % 1_000_000_000
    ev["polarity"] = rng.integers(0, 2, n)
    return ev


def _same(a, b):
    b = np.array(b, dtype=synth.EVENT_DTYPE) if not isinstance(b, np.ndarray) else b
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("x", "y", "sec", "nsec"))


@pytest.mark.parametrize("rate,msg,per_packet", [(1, 1500, 2000), (3, 977, 1500), (2, 4096, 600)])
def test_packets_match_per_event_restatement(rate, msg, per_packet):
    ev = _events(60000, 5)
    s = EventStream(0.01, per_packet, rate)
    o = StreamOracle(0.01, per_packet, rate)
    got = []
    for i in range(0, len(ev), msg):                       # messages of `msg` events, as the DVS driver delivers them
        chunk = ev[i:i + msg]
        s.eventsCallback(chunk)
        o.callback(chunk)
        while True:
            p = s.next_packet()
            if p is None:
                break
            got.append((p[0].copy(), p[1], p[2]))
    assert len(got) == len(o.packets) and len(got) > 10
    for (e, t, f), (eo, to, fo) in zip(got, o.packets):
        assert t == to and f == fo
        assert _same(e, np.array(eo, dtype=synth.EVENT_DTYPE))
    st = s.state()
    assert st["n_stored"] == o.total and st["n_subsets_pending"] == len(o.subsets) and st["n_ts_map"] == len(o.ts_keys)
    assert st["time_packet"] == o.time_packet
    s.close()


def test_windows_and_deletion_match_restatement():
    """Interleaved front-end packets and back-end windows: the look-up table, the 100-event back-off and the shared
    deletion of old events (indices of everything still pending shift)."""
    ev = _events(90000, 9)
    s = EventStream(0.01, 2000, 1)
    o = StreamOracle(0.01, 2000, 1)
    t0 = (int(ev["sec"][0]), int(ev["nsec"][0]))
    from oracle.pgo_py import dur, t_add
    wb, we = t_add(t0, dur(0.02)), t_add(t0, dur(0.12))
    n_win = 0
    for i in range(0, len(ev), 3000):
        chunk = ev[i:i + 3000]
        s.eventsCallback(chunk)
        o.callback(chunk)
        n_pk = 0
        while s.next_packet() is not None:
            n_pk += 1
        # a back-end window is cut once the store reaches 20 ms beyond its end
        last = (int(chunk["sec"][-1]), int(chunk["nsec"][-1]))
        if last > t_add(we, dur(0.02)):
            a = s.window_events(wb, we)
            b = o.window_events(wb, we)
            assert len(a) > 1000 and _same(a, np.array(b, dtype=synth.EVENT_DTYPE)), n_win
            ta = a["sec"].astype(np.int64) * 1_000_000_000 + a["nsec"]
            assert ta[-1] <= we[0] * 1_000_000_000 + we[1] - 1000          # the cut stays 1 us short of the window end
            st = s.state()
            assert st["n_stored"] == o.total and st["n_ts_map"] == len(o.ts_keys) and st["n_subsets_pending"] == len(o.subsets)
            wb, we = t_add(wb, dur(0.05)), t_add(we, dur(0.05))
            n_win += 1
    assert n_win >= 4
    assert s.state()["n_stored"] < 60000                     # old events really were deleted
    s.close()


def test_gap_marks_packet_and_uncovered_window_is_an_error():
    from cmax_slam_b200._capi import CmaxbError
    ev = _events(20000, 3, gap_at=9000, gap_s=0.5)          # half a second of silence inside one packet
    s = EventStream(0.01, 2000, 1)
    o = StreamOracle(0.01, 2000, 1)
    s.eventsCallback(ev)
    o.callback(ev)
    flags = []
    while True:
        p = s.next_packet()
        if p is None:
            break
        flags.append(p[2])
    assert flags == [f for _, _, f in o.packets] and any(flags) and not all(flags)
    t_last = (int(ev["sec"][-1]) + 10, 0)
    n_before = s.state()
    assert s.window_events((int(ev["sec"][0]), int(ev["nsec"][0])), t_last) is None      # beyond the store: "not yet", nothing consumed
    assert s.state() == n_before
    with pytest.raises(CmaxbError):
        EventStream(0.0, 2000, 1)
    s.close()


def test_randomised_configurations_match_restatement():
    """Random dt / packet size / subsampling / message size / event rate, short streams with windows now and then."""
    from oracle.pgo_py import dur, t_add
    rng = np.random.default_rng(2024)
    for trial in range(12):
        dt = float(rng.choice([0.002, 0.005, 0.01, 0.02]))
        per_packet = int(rng.integers(40, 3000))
        rate = int(rng.integers(1, 4))
        msg = int(rng.integers(50, 5000))
        ev = _events(int(rng.integers(3000, 25000)), 100 + trial, rate_hz=float(rng.choice([2e4, 1e5, 5e5])))
        s = EventStream(dt, per_packet, rate)
        o = StreamOracle(dt, per_packet, rate)
        got = []
        for i in range(0, len(ev), msg):
            s.eventsCallback(ev[i:i + msg])
            o.callback(ev[i:i + msg])
            while True:
                p = s.next_packet()
                if p is None:
                    break
                got.append((p[0].copy(), p[1], p[2]))
        assert len(got) == len(o.packets), trial
        for (e, t, f), (eo, to, fo) in zip(got, o.packets):
            assert t == to and f == fo and _same(e, np.array(eo, dtype=synth.EVENT_DTYPE)), trial
        if len(o.ts_keys) >= 4:                                 # one window over the middle of what the table covers
            tb, te = o.ts_keys[0], o.ts_keys[len(o.ts_keys) // 2]
            a = s.window_events(tb, te)
            b = o.window_events(tb, te)
            assert _same(a, np.array(b, dtype=synth.EVENT_DTYPE)), trial
        st = s.state()
        assert st["n_stored"] == o.total and st["n_ts_map"] == len(o.ts_keys) and st["n_subsets_pending"] == len(o.subsets), trial
        s.close()


@pytest.mark.parametrize("msg,per_packet", [(1500, 2000), (4096, 600), (977, 1500)])
def test_sorted_bisection_and_borrowed_messages_give_the_same_packets(msg, per_packet):
    """cmaxb_stream_push_ex: CMAXB_PUSH_SORTED finds the packet ticks inside a time-sorted message by bisection,
    CMAXB_PUSH_BORROW references the caller's buffer in place -- packets, windows and bookkeeping must equal the
    event-by-event scan over copied messages (which test_packets_match_per_event_restatement pins to the reference)."""
    ev = _events(70000, 21)
    ref = EventStream(0.01, per_packet, 1)
    streams = {"sorted": (EventStream(0.01, per_packet, 1), EventStream.PUSH_SORTED),
               "borrow": (EventStream(0.01, per_packet, 1), EventStream.PUSH_BORROW),
               "both": (EventStream(0.01, per_packet, 1), EventStream.PUSH_SORTED | EventStream.PUSH_BORROW)}
    dur = lambda sec: (int(sec), int(round((sec - int(sec)) * 1e9)))
    def t_add(t, d):
        ns = t[1] + d[1]
        return (t[0] + d[0] + ns // 1_000_000_000, ns % 1_000_000_000)
    wb = (int(ev["sec"][0]), int(ev["nsec"][0])); we = t_add(wb, dur(0.06))
    n_pk = n_win = 0
    for i in range(0, len(ev), msg):
        chunk = ev[i:i + msg]                                  # a view of `ev`: stays valid (borrowed messages)
        k0 = ref.eventsCallback(chunk)
        for name, (s, fl) in streams.items():
            assert s.eventsCallback(chunk, fl) == k0, name
        while True:
            p = ref.next_packet()
            if p is None:
                break
            n_pk += 1
            for name, (s, fl) in streams.items():
                q = s.next_packet()
                assert q is not None and q[1] == p[1] and q[2] == p[2] and _same(q[0], p[0]), name
        last = (int(chunk["sec"][-1]), int(chunk["nsec"][-1]))
        if last > t_add(we, dur(0.02)):
            a = ref.window_events(wb, we)
            for name, (s, fl) in streams.items():
                b = s.window_events(wb, we)
                assert b is not None and _same(a, b), name
                assert s.state() == ref.state(), name
            assert streams["borrow"][0].released() > 0
            wb, we = t_add(wb, dur(0.05)), t_add(we, dur(0.05))
            n_win += 1
    assert n_pk > 10 and n_win >= 3
    for s, _ in streams.values():
        s.close()
    ref.close()
