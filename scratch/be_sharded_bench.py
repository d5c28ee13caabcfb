# torchrun script: one C4 window (1e7 events, 1280x720 pano, 64 knots) sharded by time across ranks
import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
from cmax_slam_b200.dist import ShardedEventWarper
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n_ev, K, pw, ph, seed = {"C4": (10_000_000, 64, 1280, 720, 4), "C5": (50_000_000, 256, 4096, 2048, 5)}[name]
w = synth.make_be_window(n_ev // tile, K, pw, ph, seed, order=2, n_landmarks=50000)
events = np.repeat(w.events, tile) if tile > 1 else w.events      # time order preserved (cheap scale-up, as scratch/all_configs.py)
rng = np.random.default_rng(seed)
IGp = np.abs(rng.normal(0, 0.3, (ph, pw))).astype(np.float32)
x = rng.normal(0, 0.01, 3 * (K - w.n_fixed))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
sh = ShardedEventWarper(EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, pw, ph, spline_order=2, device=lr, stream=stream.cuda_stream))
sh.set_window(events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
for _ in range(3): c, g = sh.eval(x, True)
torch.cuda.synchronize()
if world > 1: dist.barrier()
N = 20
t = time.perf_counter()
for _ in range(N): c, g = sh.eval(x, True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / N
t = time.perf_counter()
for _ in range(N): c0, _ = sh.eval(x, False)
torch.cuda.synchronize()
dt0 = (time.perf_counter() - t) / N
if rank == 0:
    print(json.dumps({"config": name, "world": world, "events": len(events), "slab": sh.slab, "f+g_us": dt * 1e6, "f+g_ev_s": len(events) / dt,
                      "value_us": dt0 * 1e6, "value_ev_s": len(events) / dt0, "contrast": c, "g0": float(g[0]), "gmax": float(np.abs(g).max())}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
