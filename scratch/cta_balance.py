# per-CTA durations of the event phases of the fused kernel against what each CTA was given (events, tile segments, in-bounds events)
import sys; sys.path.insert(0, '.')
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
np.set_printoptions(linewidth=220, precision=2, suppress=True)
pk = synth.fe_config("C2")
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut)
fe.set_packet(pk.events, pk.t_ref_sec)
w = pk.omega_true + np.array([0.2, -0.1, 0.15])
for _ in range(5): fe.eval(w, True)
fe.profile(True)
T = []
for _ in range(5):
    fe.eval(w, True); T.append(fe.cta_times())
fe.profile(False)
t = np.median(np.stack(T), axis=0)
grid = t.shape[0]; n = len(pk.events)
sc = t[:, 0]; ga = t[:, 3] - t[:, 2]
# what each CTA walks: events are binned by 32x32 tile in tile order; CTA c owns records [c * chunk, (c + 1) * chunk)
tile = (pk.events["y"].astype(np.int64) // 32) * ((pk.width + 31) // 32) + pk.events["x"].astype(np.int64) // 32
counts = np.bincount(tile, minlength=((pk.width + 31) // 32) * ((pk.height + 31) // 32))
ends = np.cumsum(counts); chunk = (n + grid - 1) // grid
segs = np.array([np.searchsorted(ends, min((c + 1) * chunk, n) - 1, side="right") - np.searchsorted(ends, c * chunk, side="right") + 1 for c in range(grid)])
cells = fe.warped_cells(w)
order = np.argsort(tile, kind="stable")
inb = np.array([(cells[order[c * chunk:(c + 1) * chunk]] >= 0).mean() if c * chunk < n else 0 for c in range(grid)])
print("grid", grid, "chunk", chunk, "scatter us: min %.1f median %.1f max %.1f | gather us: min %.1f median %.1f max %.1f" % (sc.min(), np.median(sc), sc.max(), ga.min(), np.median(ga), ga.max()))
print("gather end (since entry): min %.1f median %.1f max %.1f; gather start min %.1f max %.1f" % (t[:, 3].min(), np.median(t[:, 3]), t[:, 3].max(), t[:, 2].min(), t[:, 2].max()))
for k in (1, 2, 3, 4):
    m = segs == k
    if m.any(): print("segments %d: %3d CTAs  scatter %.1f  gather %.1f  in-bounds %.2f" % (k, m.sum(), sc[m].mean(), ga[m].mean(), inb[m].mean()))
print("corr(gather, segments) %.2f  corr(gather, in-bounds) %.2f  corr(scatter, segments) %.2f  corr(scatter, in-bounds) %.2f" % (
    np.corrcoef(ga, segs)[0, 1], np.corrcoef(ga, inb)[0, 1], np.corrcoef(sc, segs)[0, 1], np.corrcoef(sc, inb)[0, 1]))
worst = np.argsort(-ga)[:12]
print("slowest gather CTAs:", [(int(c), round(float(ga[c]), 1), int(segs[c]), round(float(inb[c]), 2), round(float(t[c, 2]), 1)) for c in worst], "(cta, gather us, segments, in-bounds, gather start)")
best = np.argsort(ga)[:8]
print("fastest gather CTAs:", [(int(c), round(float(ga[c]), 1), int(segs[c]), round(float(inb[c]), 2)) for c in best])
# by SM position? CTA index modulo 148
print("gather by cta %% 3 (3 CTAs per SM, launch order):", [round(float(ga[i::3].mean()), 2) for i in range(3)], " first 148 / next 148 / last 148:", [round(float(ga[i * 148:(i + 1) * 148].mean()), 2) for i in range(3)])
# linear model of the gather / scatter time per CTA
W_, H_, r = pk.width, pk.height, 4
cx, cy = cells % W_, cells // W_
border = (cells >= 0) & ((cx <= r) | (cx >= W_ - 2 - r) | (cy <= r) | (cy >= H_ - 2 - r))
nb_ = np.array([border[order[c * chunk:(c + 1) * chunk]].sum() for c in range(grid)], float)
ni_ = np.array([(cells[order[c * chunk:(c + 1) * chunk]] >= 0).sum() for c in range(grid)], float)
ne_ = np.array([len(order[c * chunk:(c + 1) * chunk]) for c in range(grid)], float)
trow = np.array([tile[order[min(c * chunk, n - 1)]] // ((W_ + 31) // 32) for c in range(grid)])
for name, y in (("gather", ga), ("scatter", sc)):
    A = np.stack([np.ones(grid), ne_ / 1000, ni_ / 1000, segs.astype(float), nb_ / 1000], 1)
    coef, res, *_ = np.linalg.lstsq(A, y, rcond=None)
    pred = A @ coef
    print(name, "fit: const %.2f us + %.2f us/1k events + %.2f us/1k in-bounds + %.2f us/segment + %.2f us/1k border events; residual rms %.2f us (raw std %.2f)" % (*coef, np.sqrt(((y - pred) ** 2).mean()), y.std()))
print("gather by tile row of the CTA's first record:", [(int(rw), round(float(ga[trow == rw].mean()), 1)) for rw in np.unique(trow)])
print("border events per CTA: max %d, CTAs with > 100: %d" % (nb_.max(), (nb_ > 100).sum()))
