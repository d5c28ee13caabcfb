# torchrun probe: one back-end window sharded by time; wall time and per-kernel-class times of the three exchanges (rank 0)
import os, sys, time, json
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
from cmax_slam_b200.dist import ShardedEventWarper
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "C5"
w = synth.be_config(name, device=str(dev))
rng = np.random.default_rng(5)
IGp = np.abs(rng.normal(0, 0.3, (w.pano_height, w.pano_width))).astype(np.float32)
x = rng.normal(0, 0.01, 3 * (len(w.knots_xyzw) - w.n_fixed))
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
sh = ShardedEventWarper(EventWarperCMax(w.sensor_width, w.sensor_height, w.lut, w.pano_width, w.pano_height, spline_order=2, device=lr, stream=stream.cuda_stream))
sh.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
sh.eval(x, True)
sh.connect()
for mode in ("plane", "auto", "p2p"):
    sh.mode = mode
    for grad in (True, False):
        for _ in range(3): c, g = sh.eval(x, grad)
        torch.cuda.synchronize(); dist.barrier()
        t = time.perf_counter()
        for _ in range(10): c, g = sh.eval(x, grad)
        torch.cuda.synchronize()
        dt = torch.tensor([(time.perf_counter() - t) / 10], device=dev); dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0: print(name, world, mode, "f+g" if grad else "value", round(dt.item() * 1e6, 1), "us  contrast", c)
    sh.w.profile(True)
    for _ in range(4): sh.eval(x, True)
    kt = sh.w.kernel_times()
    sh.w.profile(False)
    dist.barrier()
    if rank == 0: print("   kernel classes (us per evaluation):", {k: round(v[0] / 4 * 1e3, 1) for k, v in kt.items()})
st = torch.tensor(list(sh.w.exchange_stats()), device=dev); allst = [torch.zeros_like(st) for _ in range(world)]
dist.all_gather(allst, st)
if rank == 0: print("   dirty tiles per rank:", [int(t[0]) for t in allst], "of", int(st[1]))
dist.barrier(); sh.w.exchange_close(); dist.destroy_process_group()
