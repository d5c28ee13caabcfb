"""Front-end launch ring (up to 8 evaluations queued before the first fetch), the fused kernel's variants (TMA / binning on and off),
and the fused result exchange over peer memory (two processes sharing cuda:0, CUDA IPC) -- through the C ABI,
against the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from cmax_slam_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-5


def _mk(pk, **kw):
    from cmax_slam_b200.frontend import AngVelEstimatorCMax
    fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, **kw)
    fe.set_packet(pk.events, pk.t_ref_sec)
    return fe


def test_launch_ring_fifo_matches_oracle(oracle):
    from cmax_slam_b200._capi import CmaxbError
    pk = synth.fe_config("C1", scale=0.2)
    fe = _mk(pk, max_hypotheses=3)
    oms = synth.fe_hypotheses(pk, 12, sigma=0.3)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(a, oms, True, n_threads=4)
    # 8 launches of different shapes queued back to back (they spread over the throughput lanes): k=1 f+g, k=3 value, k=2 f+g, k=1 value, ...
    plan = [(slice(0, 1), True), (slice(1, 4), False), (slice(4, 6), True), (slice(6, 7), False),
            (slice(7, 10), True), (slice(10, 11), False), (slice(11, 12), True), (slice(2, 4), True)]
    for sl, wg in plan:
        fe.eval_launch(oms[sl], wg)
    with pytest.raises(CmaxbError) as e:                      # ring full
        fe.eval_launch(oms[:1], True)
    assert e.value.code == -6
    with pytest.raises(CmaxbError) as e:                      # a synchronous evaluation would return someone else's rows
        fe.eval(oms[0], True)
    assert e.value.code == -6
    for sl, wg in plan:
        c, g = fe.eval_fetch()
        assert np.abs(c - co[sl]).max() <= RTOL * np.abs(co).max()
        if wg:
            assert np.abs(g - go[sl]).max() <= RTOL * np.abs(go).max()
    with pytest.raises(CmaxbError) as e:                      # nothing left
        fe.eval_fetch()
    assert e.value.code == -6
    # steady pipelining at depth 2 (result i-1 is read while evaluation i runs), a new packet in between
    res = []
    for lo, hi in ((0, 6), (6, 12)):
        fe.eval_launch(oms[lo], True)
        for i in range(lo + 1, hi):
            fe.eval_launch(oms[i], True)
            res.append(fe.eval_fetch())
        res.append(fe.eval_fetch())
        fe.set_packet(pk.events, pk.t_ref_sec)
    assert len(res) == 12
    for i, (c, g) in enumerate(res):
        assert abs(c[0] - co[i]) <= RTOL * co[i]
        assert np.abs(g[0] - go[i]).max() <= RTOL * np.abs(go).max()
    # set_packet with an evaluation still outstanding: the stream is drained, the result stays fetchable
    fe.eval_launch(oms[3], True)
    fe.set_packet(pk.events, pk.t_ref_sec)
    c, g = fe.eval_fetch()
    assert abs(c[0] - co[3]) <= RTOL * co[3]
    fe.close()


_GATHER_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
pk = synth.fe_config("C1", scale=0.5)
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, max_hypotheses=6)
fe.set_packet(pk.events, pk.t_ref_sec)
oms = np.concatenate([synth.fe_hypotheses(pk, 5, sigma=0.4), pk.omega_true[None, :]])
c, g = fe.eval_batch(oms, True)
c1, g1 = fe.eval(oms[5], True)
np.save(sys.argv[2], np.concatenate([c, g.ravel(), [c1], g1]))
'''


def test_fused_kernel_variants_meet_the_bar(oracle, tmp_path):
    """The fused evaluation with TMA-staged tiles + binned event walk (default), without TMA (arrival-order fallback),
    without binning, and on one lane, against the oracle; the true angular velocity (gradient near its zero crossing)
    is among the hypotheses."""
    pk = synth.fe_config("C1", scale=0.5)
    oms = np.concatenate([synth.fe_hypotheses(pk, 5, sigma=0.4), pk.omega_true[None, :]])
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(a, oms, True, n_threads=4)
    script = tmp_path / "g.py"
    script.write_text(_GATHER_WORKER)
    out = {}
    for tag, env in (("default", {}), ("no_tma", {"CMAXB_FE_TMA": "0"}), ("no_binning", {"CMAXB_FE_NO_BINNING": "1"}),
                     ("one_lane", {"CMAXB_FE_LANES": "1"})):
        path = tmp_path / f"{tag}.npy"
        r = subprocess.run([sys.executable, str(script), ROOT, str(path)], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        v = np.load(path)
        c, g, c1, g1 = v[:6], v[6:24].reshape(6, 3), v[24], v[25:]
        assert np.abs(c - co).max() <= RTOL * np.abs(co).max()
        gmax = np.abs(go).max(axis=1, keepdims=True)
        assert (np.abs(g - go) <= RTOL * gmax + 1e-7 * np.abs(go).max()).all(), (tag, np.abs(g - go).max(), gmax.ravel())
        assert abs(c1 - co[5]) <= RTOL * co[5] and np.abs(g1 - go[5]).max() <= RTOL * gmax[5] + 1e-7 * np.abs(go).max()
        out[tag] = g
    # the variants agree far below the bar
    assert np.abs(out["default"] - out["no_tma"]).max() <= 2e-6 * np.abs(go).max()
    assert np.abs(out["default"] - out["no_binning"]).max() <= 2e-6 * np.abs(go).max()


_XCHG_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.frontend import AngVelEstimatorCMax
rank, world = int(sys.argv[3]), int(sys.argv[5])
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=world)
torch.cuda.set_device(0)
pk = synth.fe_config("C1", scale=0.1)
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, max_hypotheses=2)
fe.set_packet(pk.events, pk.t_ref_sec)
oms = synth.fe_hypotheses(pk, 2 * world * 3, sigma=0.3).reshape(3, world, 2, 3)
gathered = torch.zeros(world, 2, 4, dtype=torch.float64, device="cuda:0")
fe.exchange_connect(gathered_dev_ptr=gathered.data_ptr())
out = []
# step 0: f+g, k=2, synchronous; steps 1-2: pipelined (two launches queued), value only then f+g
fe.eval_launch(oms[0, rank], True)
rows = fe.eval_fetch_all()
torch.cuda.synchronize()
assert np.array_equal(rows, gathered.cpu().numpy())
out.append(rows)
fe.eval_launch(oms[1, rank], False)
fe.eval_launch(oms[2, rank], True)
out.append(fe.eval_fetch_all())
out.append(fe.eval_fetch_all())
dist.barrier()
fe.exchange_close()
c, g = fe.eval_batch(oms[0, rank], True)          # exchange off again: plain evaluation still works
assert np.allclose(c, out[0][rank, :, 0], rtol=1e-6)
np.save(sys.argv[4] + f"/x{rank}.npy", np.stack(out))
dist.barrier()
fe.close()
dist.destroy_process_group()
print("ok")
'''


@pytest.mark.parametrize("world", [2, 3])
def test_fused_result_exchange_between_processes(oracle, tmp_path, world):
    import socket
    script = tmp_path / "x.py"
    script.write_text(_XCHG_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path), str(world)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=240)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append("timeout: " + p.communicate()[0])
    assert all(p.returncode == 0 for p in procs), outs
    got = [np.load(tmp_path / f"x{r}.npy") for r in range(world)]      # [3 steps][world][2][4]
    for r in range(1, world):
        assert np.array_equal(got[0], got[r])                              # every rank holds the same gathered rows
    pk = synth.fe_config("C1", scale=0.1)
    oms = synth.fe_hypotheses(pk, 2 * world * 3, sigma=0.3).reshape(3, world, 2, 3)
    a = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(a, oms.reshape(-1, 3), True, n_threads=4)
    co, go = co.reshape(3, world, 2), go.reshape(3, world, 2, 3)
    rows = got[0]
    assert np.abs(rows[..., 0] - co).max() <= RTOL * np.abs(co).max()
    assert np.abs(rows[1][..., 1:]).max() == 0.0                            # value-only step: zero gradient columns
    for st in (0, 2):
        assert np.abs(rows[st][..., 1:] - go[st]).max() <= RTOL * np.abs(go).max()


_C3_WORKER = r'''
import sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from cmax_slam_b200 import synth
from cmax_slam_b200.dist import sharded_eval_fused
from cmax_slam_b200.frontend import AngVelEstimatorCMax
rank, world = int(sys.argv[3]), 2
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=rank, world_size=world)
torch.cuda.set_device(0)
pk = synth.fe_config("C1", scale=0.1)
fe = AngVelEstimatorCMax(pk.width, pk.height, pk.K, pk.lut, max_hypotheses=8)
fe.set_packet(pk.events, pk.t_ref_sec)
fe.exchange_connect()
oms = synth.fe_hypotheses(pk, 16, sigma=0.3)
c, g = sharded_eval_fused(fe, oms, True)
cv, gv = sharded_eval_fused(fe, oms[:6], False)
assert gv is None
np.save(sys.argv[4] + f"/c3_{rank}.npy", np.concatenate([c, g.ravel(), cv]))
dist.barrier()
fe.exchange_close()
fe.close()
dist.destroy_process_group()
print("ok")
'''


def test_c3_style_hypothesis_sweep_with_fused_exchange(oracle, tmp_path):
    """16 hypotheses over 2 ranks (8 per launch), rows re-interleaved: equal on both ranks and to the oracle."""
    import socket
    script = tmp_path / "c3.py"
    script.write_text(_C3_WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r), str(tmp_path)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=240)[0])
        except subprocess.TimeoutExpired:
            p.kill()
            outs.append("timeout: " + p.communicate()[0])
    assert all(p.returncode == 0 for p in procs), outs
    a, b = np.load(tmp_path / "c3_0.npy"), np.load(tmp_path / "c3_1.npy")
    assert np.array_equal(a, b)
    pk = synth.fe_config("C1", scale=0.1)
    oms = synth.fe_hypotheses(pk, 16, sigma=0.3)
    args = oracle.fe_args(pk.events, pk.t_ref_sec, pk.lut, pk.width, pk.height, pk.K)
    co, go = oracle.fe_eval_batch(args, oms, True, n_threads=4)
    c, g, cv = a[:16], a[16:64].reshape(16, 3), a[64:]
    assert np.abs(c - co).max() <= RTOL * np.abs(co).max()
    assert np.abs(g - go).max() <= RTOL * np.abs(go).max()
    assert np.abs(cv - co[:6]).max() <= RTOL * np.abs(co).max()
