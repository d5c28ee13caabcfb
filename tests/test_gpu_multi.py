"""Multi-GPU paths over REAL peers (skipped on a box with one GPU; the same logic runs there as two processes on one device,
tests/test_gpu_sharded.py and tests/test_gpu_fe_pipeline.py): the back end's three exchanges of a time-sharded window (NCCL
all-reduce of the plane, NCCL reduce-scatter / all-gather by row band, kernels over peer memory) against the un-sharded
evaluation, one process per GPU under torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.backend import EventWarperCMax
from cmax_slam_b200.dist import ShardedEventWarper
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
K_T = (60.0, 61.0, 31.5, 23.5)
w = synth.make_be_window(40001, 10, 256, 128, 43, order=2, sensor=(64, 48), K4=K_T, n_landmarks=300, n_fixed=1)
rng = np.random.default_rng(3)
IGp = np.abs(rng.normal(0, 0.3, (128, 256))).astype(np.float32)
x = rng.normal(0, 0.02, 27)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ref = EventWarperCMax(64, 48, w.lut, 256, 128, spline_order=2, device=lr)
ref.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
c0, g0 = ref.eval(x, True)
sh = ShardedEventWarper(EventWarperCMax(64, 48, w.lut, 256, 128, spline_order=2, device=lr, stream=stream.cuda_stream))
sh.set_window(w.events, w.knots_xyzw, w.t0_ns, w.dt_ns, w.n_fixed, w.tnext, IGp, 0.5)
sh.connect()
for mode in ("plane", "bands", "p2p"):
    sh.mode = mode
    for xx, wg in ((x, True), (x, False), (0.5 * x, True), (x, True)):
        c, g = sh.eval(xx, wg)
    assert abs(c - c0) <= 1e-6 * abs(c0), (mode, c, c0)
    assert np.abs(g - g0).max() <= 1e-6 * np.abs(g0).max(), (mode, np.abs(g - g0).max())
    t = torch.tensor(np.concatenate([[c], g]), device=dev)
    all_t = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_t, t)
    if mode == "p2p":
        assert all(torch.equal(all_t[0], a) for a in all_t), "p2p results differ between ranks"      # sums formed in rank order
    else:
        assert all(torch.allclose(all_t[0], a, rtol=1e-12, atol=0) for a in all_t)
dirty, total = sh.w.exchange_stats()
assert 0 < dirty <= total
dist.barrier()
sh.w.exchange_close()
dist.destroy_process_group()
print("ok", rank)
'''


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_back_end_exchanges_across_two_gpus(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    n = min(_n_gpus(), 4)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script), ROOT], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.count("ok ") == n, p.stdout[-3000:]
