"""TEST INFRASTRUCTURE.  Line-by-line Python restatement of the reference's event ingestion -- one event at a time,
exactly as AngVelEstimator::pushEvent does (src/frontend/ang_vel_estimator.cpp:68-183), with
PoseGraphOptimizer::getEventSubset (src/backend/pose_graph_optimizer.cpp:133-166) -- used to check csrc/stream.cu.
Pure-Python loops: small cases only.  The reference solves a packet inside pushEvent; `on_packet` stands in for that."""
import bisect

from .pgo_py import dur, t_add


class StreamOracle:
    def __init__(self, dt_ang_vel, num_events_per_packet, sample_rate=1, on_packet=None):
        self.dt = dt_ang_vel
        self.dt_av = dur(dt_ang_vel)
        self.half = num_events_per_packet // 2
        self.rate = sample_rate
        self.on_packet = on_packet
        self.events = []            # (x, y, sec, nsec) tuples
        self.total = 0
        self.init = False
        self.subsets = []
        self.ts_keys, self.ts_vals = [], []       # ev_subset_ts_map_
        self.ev_beg = self.ev_end = 0
        self.packets = []

    def callback(self, msg):
        for i in range(0, len(msg), self.rate):
            self.push(msg[i])

    def push(self, e):
        ts = (int(e["sec"]), int(e["nsec"]))
        if not self.init:
            self.time_packet = t_add(ts, dur((self.dt_av[0] + 1e-9 * self.dt_av[1]) * 0.5))
            self.time_get_subset = self.time_packet
            self.init = True
        self.events.append(e)
        self.total += 1
        if ts > self.time_get_subset:
            self.subsets.append((max(self.total - self.half, 0), self.total + self.half))
            i = bisect.bisect_left(self.ts_keys, ts)
            if not (i < len(self.ts_keys) and self.ts_keys[i] == ts):
                self.ts_keys.insert(i, ts); self.ts_vals.insert(i, self.total - 1)
            self.time_get_subset = t_add(self.time_get_subset, self.dt_av)
        if self.subsets and self.total > self.subsets[0][1]:
            self.ev_beg, self.ev_end = self.subsets.pop(0)
            sub = self.events[self.ev_beg:self.ev_end]
            f, l = sub[0], sub[-1]
            ds, dn = int(l["sec"]) - int(f["sec"]), int(l["nsec"]) - int(f["nsec"])
            if dn < 0:
                dn += 1000000000; ds -= 1
            too_long = (ds + 1e-9 * dn) > 10 * self.dt
            self.packets.append((sub, self.time_packet, too_long))
            if self.on_packet:
                self.on_packet(sub, self.time_packet, too_long)
            self.time_packet = t_add(self.time_packet, self.dt_av)

    def delete_old(self, idx_backend):
        n = min(idx_backend, self.ev_beg)
        if n <= 0:
            return
        del self.events[:n]
        self.total -= n
        self.ev_beg -= n; self.ev_end -= n
        self.subsets = [(a - n, b - n) for a, b in self.subsets]
        self.ts_vals = [v - n for v in self.ts_vals]

    def window_events(self, t_beg, t_end):
        ib = bisect.bisect_right(self.ts_keys, t_beg)
        ie = bisect.bisect_left(self.ts_keys, t_end)
        if ib >= len(self.ts_keys) or ie >= len(self.ts_keys):
            raise IndexError("store does not cover the window")
        beg, end = self.ts_vals[ib], self.ts_vals[ie]
        t_end_mod = t_add(t_end, (0, -1000))
        while (int(self.events[end]["sec"]), int(self.events[end]["nsec"])) > t_end_mod:
            end -= 100
            if end <= beg:
                end = beg + 1
                break
        sub = self.events[beg:end]
        del self.ts_keys[:ib + 1]; del self.ts_vals[:ib + 1]
        self.delete_old(beg)
        return sub
