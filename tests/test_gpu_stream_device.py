"""Device-resident event store (cmaxb_stream_attach_device / _push_ex / _next_packet_device, csrc/stream.cu): the packets
handed out as views of the device ring must be the packets the host path hands out, and evaluating them through
cmaxb_fe_set_packet_view must give the result of uploading the host packet."""
import numpy as np
import pytest

from cmax_slam_b200 import synth
from cmax_slam_b200.stream import EventStream

pytestmark = pytest.mark.gpu


def _stream_events(n, seed, W, H, rate_hz=4.0e5):
    rng = np.random.default_rng(seed)
    dt = rng.exponential(1.0 / rate_hz, n)
    t_ns = (synth.EPOCH_SEC * 1_000_000_000 + 123_456_789 + np.cumsum(dt * 1e9)).astype(np.int64)
    ev = np.zeros(n, synth.EVENT_DTYPE)
    ev["x"] = rng.integers(0, W, n); ev["y"] = rng.integers(0, H, n)
    ev["sec"] = t_ns // 1_000_000_000; ev["nsec"] = t_ns % 1_000_000_000
    return ev


@pytest.mark.parametrize("flags", [0, EventStream.PUSH_SORTED | EventStream.PUSH_BORROW])
def test_device_packets_equal_host_packets(flags):
    import torch
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    W, H, K4 = 240, 180, synth.K_SMALL
    per_packet, msg = 6000, 2500
    n = 90000
    pinned = torch.empty(n * 16, dtype=torch.uint8).pin_memory()
    ev = pinned.numpy().view(synth.EVENT_DTYPE)
    ev[:] = _stream_events(n, 3, W, H)
    host = EventStream(0.01, per_packet, 1)
    dev = EventStream(0.01, per_packet, 1)
    fe = AngVelEstimatorCMax(W, H, K4, synth.bearing_lut(W, H, K4), packet_slots=2)
    stream = torch.cuda.Stream()
    dev.attach_device(0, stream.cuda_stream, ring_events=4 * per_packet)      # small ring: wraps several times
    w = np.array([0.5, -0.8, 1.5])
    n_pk = 0
    for i in range(0, n, msg):
        chunk = ev[i:i + msg]
        k0 = host.eventsCallback(chunk)
        assert dev.eventsCallback(chunk, flags) == k0
        while True:
            p = host.next_packet()
            if p is None:
                break
            q = dev.next_packet_device()
            assert q is not None and q[1] == p[1] and q[2] == p[2] and q[0][1] == len(p[0])
            dev.wait_copied(0)            # legacy default stream waits for the store's copy stream
            stream.synchronize()
            got = torch.empty(len(p[0]) * 16, dtype=torch.uint8, device="cuda")
            import ctypes
            from cmax_slam_b200 import _capi  # noqa: F401
            # copy the ring view out with torch (device to device through the raw pointer)
            class _View:
                __cuda_array_interface__ = {"shape": (len(p[0]) * 16,), "typestr": "|u1", "data": (q[0][0], False), "version": 2}
            got.copy_(torch.as_tensor(_View(), device="cuda"))
            assert np.array_equal(got.cpu().numpy().view(synth.EVENT_DTYPE), p[0])
            # evaluation of the view == evaluation of the uploaded host packet
            t_ref = float(p[1][0]) + 1e-9 * float(p[1][1])
            fe.select_packet(0); fe.set_packet(p[0], t_ref)
            c0, g0 = fe.eval(w, True)
            fe.select_packet(1); fe.set_packet(q[0], t_ref, view=True)
            c1, g1 = fe.eval(w, True)
            assert abs(c1 - c0) <= 1e-6 * abs(c0) and np.abs(g1 - g0).max() <= 1e-5 * np.abs(g0).max()
            n_pk += 1
    assert n_pk > 10
    fe.close(); host.close(); dev.close()
