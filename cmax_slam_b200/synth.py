"""Synthetic event packets / windows for the configurations of BASELINE.json.

All draws use numpy.random.Generator(PCG64(seed)); events are sorted by time and carry integer
microsecond timestamps on an absolute epoch (like real `dvs_msgs::Event` stamps), so the f64
`toSec()` quantisation of the reference (local_image_warped_events.cpp:75) is exercised.

Record layout = `dvs_msgs::Event` as generated for C++: uint16 x, uint16 y, ros::Time {u32 sec,
u32 nsec}, bool polarity (+3 pad) -> 16 bytes (see include/cmax_b200.h `cmaxb_event`).
"""
from dataclasses import dataclass, field

import numpy as np

EVENT_DTYPE = np.dtype(
    [("x", "<u2"), ("y", "<u2"), ("sec", "<u4"), ("nsec", "<u4"), ("polarity", "u1"), ("pad", "u1", (3,))]
)
EPOCH_SEC = 1_600_000_000  # absolute stamps of the order of real recordings

# launch/ecrot_handheld.launch:50 intrinsics (640x480)
K_ECROT = (588.0999, 593.9887, 339.8259, 242.4252)
K_SMALL = (200.0, 200.0, 120.0, 90.0)


def bearing_lut(W, H, K4):
    """Ideal-pinhole stand-in for CMaxSLAM::precomputeBearingVectors (src/cmax_slam.cpp:106-120):
    ((x-cx)/fx, (y-cy)/fy, 1) per pixel, row-major, float64 [H*W,3]."""
    fx, fy, cx, cy = K4
    xs = (np.arange(W, dtype=np.float64) - cx) / fx
    ys = (np.arange(H, dtype=np.float64) - cy) / fy
    lut = np.empty((H, W, 3), np.float64)
    lut[..., 0] = xs[None, :]
    lut[..., 1] = ys[:, None]
    lut[..., 2] = 1.0
    return lut.reshape(-1, 3)


def _pack_events(x, y, t_us_rel, pol=None):
    """t_us_rel: int64 microseconds relative to EPOCH_SEC."""
    order = np.argsort(t_us_rel, kind="stable")
    x, y, t = x[order], y[order], t_us_rel[order]
    ev = np.zeros(len(x), EVENT_DTYPE)
    ev["x"] = x
    ev["y"] = y
    ev["sec"] = EPOCH_SEC + t // 1_000_000
    ev["nsec"] = (t % 1_000_000) * 1000
    ev["polarity"] = (np.arange(len(x)) & 1) if pol is None else pol[order]
    return ev


@dataclass
class FePacket:
    events: np.ndarray
    t_ref_sec: float
    lut: np.ndarray
    width: int
    height: int
    K: tuple
    omega_true: np.ndarray
    batch_size: int = 100
    blur_sigma: float = 1.0
    name: str = ""


def make_fe_packet(n_events, W, H, K4, seed, n_landmarks, omega_true=(0.6, -1.1, 2.3),
                   half_span_s=0.025, name=""):
    """Front-end packet (configs C1/C2/C3): landmarks uniform on the sensor at t_ref; an event of
    landmark L at time t fires at pixel round(project((I - [w*(t - t_ref)]x) L))."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fx, fy, cx, cy = K4
    w = np.asarray(omega_true, np.float64)
    lx = rng.uniform(0, W - 1, n_landmarks)
    ly = rng.uniform(0, H - 1, n_landmarks)
    L = np.stack([(lx - cx) / fx, (ly - cy) / fy, np.ones(n_landmarks)], 1)
    t_ref_us = 500_000  # relative to the epoch second
    span_us = int(round(half_span_s * 1e6))
    xs, ys, ts = [], [], []
    need = n_events
    while need > 0:
        m = int(need * 1.3) + 1024
        li = rng.integers(0, n_landmarks, m)
        t_us = rng.integers(t_ref_us - span_us, t_ref_us + span_us + 1, m)
        dt = (t_us - t_ref_us) * 1e-6
        b = L[li]
        rot = b - np.cross(w[None, :] * dt[:, None], b)
        px = np.rint(fx * rot[:, 0] / rot[:, 2] + cx).astype(np.int64)
        py = np.rint(fy * rot[:, 1] / rot[:, 2] + cy).astype(np.int64)
        ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
        xs.append(px[ok][:need]); ys.append(py[ok][:need]); ts.append(t_us[ok][:need])
        need -= len(xs[-1])
    x = np.concatenate(xs); y = np.concatenate(ys); t = np.concatenate(ts)
    ev = _pack_events(x, y, t)
    t_ref_sec = float(EPOCH_SEC) + 1e-9 * float(t_ref_us * 1000)  # == ros::Time(sec,nsec).toSec()
    return FePacket(ev, t_ref_sec, bearing_lut(W, H, K4), W, H, tuple(K4), w, name=name)


def make_fe_stream(n_events, W, H, K4, seed, rate_hz=2.0e7, n_landmarks=60000, amp_rad=0.08, period_s=0.2,
                   axis=(0.6, -1.1, 2.3)):
    """A continuous front-end event stream (what the DVS driver delivers message by message): `n_events` events at a
    constant rate, time-sorted with nanosecond stamps; the camera oscillates about `axis` with amplitude `amp_rad`
    (angle = amp sin(2 pi t / period)), landmarks lie on the sensor plane extended by the largest displacement.
    Returns (events, lut).  rate 2e7 events/s = 200 000 events per 10 ms angular-velocity tick, i.e. the C2 packet
    (1 000 000 events over 50 ms) as `num_events_per_packet` of the reference's packet cutter."""
    rng = np.random.Generator(np.random.PCG64(seed))
    fx, fy, cx, cy = K4
    ax = np.asarray(axis, np.float64); ax = ax / np.linalg.norm(ax)
    margin = int(np.ceil(amp_rad * max(fx, fy) * 1.6)) + 8
    lx = rng.uniform(-margin, W - 1 + margin, n_landmarks)
    ly = rng.uniform(-margin, H - 1 + margin, n_landmarks)
    L = np.stack([(lx - cx) / fx, (ly - cy) / fy, np.ones(n_landmarks)], 1)
    xs, ys, ts = [], [], []
    need = n_events
    t_cursor = 0.0
    chunk = 2_000_000
    while need > 0:
        m = min(chunk, int(need * 1.8) + 4096)
        span = m / (rate_hz * 1.6)                    # ~60 % of the candidates survive: keep the event RATE at rate_hz
        t = np.sort(rng.uniform(t_cursor, t_cursor + span, m))
        li = rng.integers(0, n_landmarks, m)
        th = amp_rad * np.sin(2 * np.pi * t / period_s)
        b = L[li]
        k = ax[None, :]
        c, sn = np.cos(th)[:, None], np.sin(th)[:, None]
        rot = b * c + np.cross(k, b) * sn + k * (b @ ax)[:, None] * (1 - c)     # Rodrigues
        px = np.rint(fx * rot[:, 0] / rot[:, 2] + cx).astype(np.int64)
        py = np.rint(fy * rot[:, 1] / rot[:, 2] + cy).astype(np.int64)
        ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
        idx = np.nonzero(ok)[0][:need]
        xs.append(px[idx]); ys.append(py[idx]); ts.append(t[idx])
        need -= len(idx)
        t_cursor = t[idx[-1]] if len(idx) else t_cursor + span
    x = np.concatenate(xs); y = np.concatenate(ys); t = np.concatenate(ts)
    t_ns = (t * 1e9).astype(np.int64) + 500_000_000
    ev = np.zeros(len(x), EVENT_DTYPE)
    ev["x"] = x; ev["y"] = y
    ev["sec"] = EPOCH_SEC + t_ns // 1_000_000_000
    ev["nsec"] = t_ns % 1_000_000_000
    ev["polarity"] = np.arange(len(x)) & 1
    return ev, bearing_lut(W, H, K4)


def fe_config(name, scale=1.0):
    """BASELINE.json configs: 'C1' (1e5 ev, 240x180), 'C2' (1e6 ev, 640x480). `scale` shrinks the
    event count for tests."""
    if name == "C1":
        return make_fe_packet(int(100_000 * scale), 240, 180, K_SMALL, 1, 2000, name="C1")
    if name in ("C2", "C3"):
        return make_fe_packet(int(1_000_000 * scale), 640, 480, K_ECROT, 2 if name == "C2" else 3,
                              20000, name=name)
    raise ValueError(name)


def fe_hypotheses(pkt, k, seed=3, sigma=0.5):
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    return pkt.omega_true[None, :] + rng.normal(0, sigma, (k, 3))


# ------------------------------------------------------------------------------------------------
# Back-end windows (C4 / C5)
# ------------------------------------------------------------------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def _qexp(w):
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    small = th < 1e-12
    s = np.where(small, 0.5, np.sin(0.5 * th) / np.where(small, 1.0, th))
    return np.concatenate([s * w, np.cos(0.5 * th)], -1)


def _qlog(q):
    n = np.linalg.norm(q[..., :3], axis=-1, keepdims=True)
    w = q[..., 3:4]
    f = np.where(n < 1e-12, 2.0 / w, 2.0 * np.arctan(n / w) / np.where(n < 1e-12, 1.0, n))
    return f * q[..., :3]


def _qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qrot(q, v):
    """Rotate v by unit quaternion q (xyzw)."""
    u = q[..., :3]
    w = q[..., 3:4]
    t = 2.0 * np.cross(u, v)
    return v + w * t + np.cross(u, t)


@dataclass
class BeWindow:
    events: np.ndarray
    lut: np.ndarray
    sensor_width: int
    sensor_height: int
    pano_width: int
    pano_height: int
    knots_xyzw: np.ndarray
    t0_ns: int
    dt_ns: int
    spline_order: int
    n_fixed: int
    tnext: tuple  # (sec, nsec)
    IGp: np.ndarray = None
    alpha: float = 0.5
    batch_size: int = 100
    event_sample_rate: int = 1
    blur_sigma: float = 1.0
    name: str = ""
    extra: dict = field(default_factory=dict)


def make_be_window(n_events, n_knots, pano_w, pano_h, seed, order=2, sensor=(640, 480), K4=K_ECROT,
                   dt_knots_s=0.05, knot_sigma=0.04, n_landmarks=50000, n_fixed=1, name=""):
    """Back-end window: knots R0 = I, R_{k+1} = R_k exp(N(0, knot_sigma^2)); landmarks on the part of
    the unit sphere the camera sweeps; event of landmark d at time t fires at round(pinhole(R(t)^T d)),
    where R(t) is the geodesic interpolation between the two nearest knots (the generator does not
    need the exact spline)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    W, H = sensor
    fx, fy, cx, cy = K4
    knots = np.zeros((n_knots, 4))
    knots[0] = [0, 0, 0, 1]
    for k in range(1, n_knots):
        knots[k] = _qmul(knots[k - 1], _qexp(rng.normal(0, knot_sigma, 3)))
        knots[k] /= np.linalg.norm(knots[k])
    dt_ns = int(round(dt_knots_s * 1e9))
    t0_rel_us = 250_000
    t0_ns = EPOCH_SEC * 1_000_000_000 + t0_rel_us * 1000
    n_seg = n_knots - order + 1          # valid span of an order-N cumulative spline
    span_us = n_seg * (dt_ns // 1000)
    # landmarks: uniform on the cap that the optical axis sweeps (+ half the diagonal FOV)
    axis = _qrot(knots, np.array([0.0, 0.0, 1.0]))
    half_fov = np.arctan(np.hypot(W / (2 * fx), H / (2 * fy)))
    ang = np.arccos(np.clip(axis[:, 2], -1, 1)).max() + half_fov + 0.05
    ang = min(ang, np.pi)
    z = rng.uniform(np.cos(ang), 1.0, n_landmarks)
    ph = rng.uniform(0, 2 * np.pi, n_landmarks)
    s = np.sqrt(1 - z * z)
    Lw = np.stack([s * np.cos(ph), s * np.sin(ph), z], 1)

    xs, ys, ts = [], [], []
    need = n_events
    chunk = 2_000_000
    while need > 0:
        m = min(chunk, int(need * 3) + 4096)
        li = rng.integers(0, n_landmarks, m)
        t_us = rng.integers(1, span_us - 1, m)        # strictly inside the valid span
        sidx = t_us // (dt_ns // 1000)
        u = (t_us % (dt_ns // 1000)) / (dt_ns / 1000.0)
        qa = knots[sidx]
        qb = knots[sidx + 1]
        d = _qlog(_qmul(_qconj(qa), qb))
        q = _qmul(qa, _qexp(d * u[:, None]))
        pc = _qrot(_qconj(q), Lw[li])                 # R(t)^T d
        zc = pc[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            px = np.rint(fx * pc[:, 0] / zc + cx)
            py = np.rint(fy * pc[:, 1] / zc + cy)
        ok = (zc > 1e-3) & (px >= 0) & (px < W) & (py >= 0) & (py < H)
        xs.append(px[ok][:need].astype(np.int64)); ys.append(py[ok][:need].astype(np.int64))
        ts.append(t_us[ok][:need])
        need -= len(xs[-1])
    x = np.concatenate(xs); y = np.concatenate(ys); t = np.concatenate(ts) + t0_rel_us
    ev = _pack_events(x, y, t)
    t_next_us = t0_rel_us + span_us // 2
    tnext = (EPOCH_SEC + t_next_us // 1_000_000, (t_next_us % 1_000_000) * 1000)
    return BeWindow(ev, bearing_lut(W, H, K4), W, H, pano_w, pano_h, knots, t0_ns, dt_ns, order,
                    n_fixed, tnext, name=name)


def make_be_window_torch(n_events, n_knots, pano_w, pano_h, seed, order=2, sensor=(640, 480), K4=K_ECROT,
                         dt_knots_s=0.05, knot_sigma=0.04, n_landmarks=50000, n_fixed=1, name="", device="cuda"):
    """make_be_window with the per-event sampling on a torch device (the numpy version needs ~5 s per million events;
    the full-size C4 / C5 windows of the benchmark are drawn on the GPU in about a second).  Same construction -- knots,
    landmark cap, geodesic interpolation, rounding to the sensor grid -- with torch's generator for the per-event draws, so
    the events differ from the numpy version's (both are synthetic; tests that need the oracle use the numpy one)."""
    import torch
    rng = np.random.Generator(np.random.PCG64(seed))
    W, H = sensor
    fx, fy, cx, cy = K4
    knots = np.zeros((n_knots, 4))
    knots[0] = [0, 0, 0, 1]
    for k in range(1, n_knots):
        knots[k] = _qmul(knots[k - 1], _qexp(rng.normal(0, knot_sigma, 3)))
        knots[k] /= np.linalg.norm(knots[k])
    dt_ns = int(round(dt_knots_s * 1e9))
    t0_rel_us = 250_000
    t0_ns = EPOCH_SEC * 1_000_000_000 + t0_rel_us * 1000
    n_seg = n_knots - order + 1
    seg_us = dt_ns // 1000
    span_us = n_seg * seg_us
    axis = _qrot(knots, np.array([0.0, 0.0, 1.0]))
    half_fov = np.arctan(np.hypot(W / (2 * fx), H / (2 * fy)))
    ang = min(np.arccos(np.clip(axis[:, 2], -1, 1)).max() + half_fov + 0.05, np.pi)
    z = rng.uniform(np.cos(ang), 1.0, n_landmarks)
    ph = rng.uniform(0, 2 * np.pi, n_landmarks)
    sq = np.sqrt(1 - z * z)
    dev = torch.device(device)
    Lw = torch.tensor(np.stack([sq * np.cos(ph), sq * np.sin(ph), z], 1), dtype=torch.float64, device=dev)
    # per-segment rotation (knot k) and relative rotation vector, as rotation matrices / vectors on the device
    qa = torch.tensor(knots[:-1], dtype=torch.float64, device=dev)
    d = torch.tensor(_qlog(_qmul(_qconj(knots[:-1]), knots[1:])), dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)

    def qmul(a, b):
        ax, ay, az, aw = a.unbind(-1)
        bx, by, bz, bw = b.unbind(-1)
        return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                            aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz], -1)

    def qexp(w):
        th = w.norm(dim=-1, keepdim=True)
        small = th < 1e-12
        sc = torch.where(small, torch.full_like(th, 0.5), torch.sin(0.5 * th) / torch.where(small, torch.ones_like(th), th))
        return torch.cat([sc * w, torch.cos(0.5 * th)], -1)

    def qrot_conj(q, v):     # rotate v by the conjugate of q
        u = -q[..., :3]
        w = q[..., 3:4]
        t = 2.0 * torch.cross(u, v, dim=-1)
        return v + w * t + torch.cross(u, t, dim=-1)

    xs, ys, ts = [], [], []
    need = n_events
    chunk = 4_000_000
    while need > 0:
        m = min(chunk, int(need * 3) + 4096)
        li = torch.randint(0, n_landmarks, (m,), generator=gen, device=dev)
        t_us = torch.randint(1, span_us - 1, (m,), generator=gen, device=dev)
        sidx = torch.div(t_us, seg_us, rounding_mode="floor")
        u = (t_us % seg_us).to(torch.float64) / float(dt_ns / 1000.0)
        q = qmul(qa[sidx], qexp(d[sidx] * u[:, None]))
        pc = qrot_conj(q, Lw[li])
        zc = pc[:, 2]
        px = torch.round(fx * pc[:, 0] / zc + cx)
        py = torch.round(fy * pc[:, 1] / zc + cy)
        ok = (zc > 1e-3) & (px >= 0) & (px < W) & (py >= 0) & (py < H)
        sel = torch.nonzero(ok).squeeze(1)[:need]
        xs.append(px[sel].to(torch.int64).cpu().numpy()); ys.append(py[sel].to(torch.int64).cpu().numpy())
        ts.append(t_us[sel].cpu().numpy())
        need -= len(xs[-1])
    x = np.concatenate(xs); y = np.concatenate(ys); t = np.concatenate(ts) + t0_rel_us
    ev = _pack_events(x, y, t)
    t_next_us = t0_rel_us + span_us // 2
    tnext = (EPOCH_SEC + t_next_us // 1_000_000, (t_next_us % 1_000_000) * 1000)
    return BeWindow(ev, bearing_lut(W, H, K4), W, H, pano_w, pano_h, knots, t0_ns, dt_ns, order,
                    n_fixed, tnext, name=name)


def be_config(name, scale=1.0, order=2, device=None):
    """'C4': 1e7 ev, 64 knots, 1280x720 pano; 'C5': 5e7 ev, 256 knots, 4096x2048 pano."""
    make = make_be_window if device is None else (lambda *a, **kw: make_be_window_torch(*a, device=device, **kw))
    if name == "C4":
        return make(int(10_000_000 * scale), 64, 1280, 720, 4, order=order,
                    n_landmarks=50000, n_fixed=1 if order == 2 else 3, name="C4")
    if name == "C5":
        return make(int(50_000_000 * scale), 256, 4096, 2048, 5, order=order,
                    n_landmarks=200000, n_fixed=1 if order == 2 else 3, name="C5")
    raise ValueError(name)
