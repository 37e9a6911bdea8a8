"""Host-side logic that needs no GPU: synthetic shapes / byte accounting,
GOP-job partitioning (8e) and the world_size-2 gloo path of the multi-GPU helpers."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT


def test_synthetic_shapes_and_bytes():
    from deepsvc_b200 import synthetic
    d = synthetic.make_pframe_inputs(B=1, H=256, W=448)
    assert [tuple(t.shape) for t in d["pyr_img"]] == [(1, 3, 32, 56), (1, 3, 64, 112), (1, 3, 128, 224), (1, 3, 256, 448)]
    assert tuple(d["feature"].shape) == (1, 64, 256, 448) and tuple(d["flow"].shape) == (1, 2, 256, 448)
    assert tuple(d["mv_y"].shape) == (1, 64, 16, 28) and tuple(d["mv_z"].shape) == (1, 64, 4, 7)
    assert tuple(d["res_y"].shape) == (1, 96, 16, 28) and tuple(d["res_z"].shape) == (1, 96, 4, 7)
    assert float(d["pyr_flow"][0].abs().max()) == 0.0
    # the per-frame algorithmic bytes quoted in SURVEY.md 8d / BASELINE.md
    mb = lambda b: round(b / 1e6, 1)
    b1 = synthetic.pframe_algorithmic_bytes(1, 256, 448)
    assert mb(b1["total"]) == 69.4
    b2 = synthetic.pframe_algorithmic_bytes(1, 1088, 1920)
    assert (mb(b2["spynet"]), mb(b2["frame"]), mb(b2["feature"]), mb(b2["entropy"]), mb(b2["total"])) == \
        (88.8, 66.8, 1086.3, 21.5, 1263.4)
    assert mb(synthetic.pframe_algorithmic_bytes(1, 2176, 3840)["total"]) == 5053.7
    # same seed -> same bits; different seed -> different
    e = synthetic.make_pframe_inputs(B=1, H=256, W=448)
    assert torch.equal(d["mv_y"], e["mv_y"]) and torch.equal(d["flow"], e["flow"])
    f = synthetic.make_pframe_inputs(B=1, H=256, W=448, seed=17)
    assert not torch.equal(d["flow"], f["flow"])
    # ties and far tails are present
    r = (d["mv_y"] - d["mv_means"])
    assert int(((r - torch.floor(r)) == 0.5).sum()) > 0


def test_gop_partitioning_matches_survey():
    from deepsvc_b200 import shard
    jobs = shard.make_gop_jobs([96] * 7, 32)          # config 4: 7 sequences x 96 frames, GOP 32
    assert len(jobs) == 21 and all(j.p_frames == 31 for j in jobs)
    for world, counts in ((8, [3, 3, 3, 3, 3, 2, 2, 2]), (4, [6, 5, 5, 5]), (2, [11, 10])):
        a = shard.assign_jobs(jobs, world)
        assert sorted((len(x) for x in a), reverse=True) == counts
        assert sorted((j.sequence, j.gop) for x in a for j in x) == sorted((j.sequence, j.gop) for j in jobs)
        assert sum(len(x) for x in a) == 21
    assert abs(shard.balance(shard.assign_jobs(jobs, 8)) - 0.875) < 1e-9
    jobs12 = shard.make_gop_jobs([96] * 7, 12)        # the reference's own GOP (test_video.py:22)
    assert len(jobs12) == 56
    assert shard.balance(shard.assign_jobs(jobs12, 8)) == 1.0
    # ragged: last GOP shorter, every job appears exactly once, deterministic
    jobs = shard.make_gop_jobs([50, 7, 33], 32)
    assert [(j.sequence, j.gop, j.first_frame, j.n_frames) for j in jobs] == \
        [(0, 0, 0, 32), (0, 1, 32, 18), (1, 0, 0, 7), (2, 0, 0, 32), (2, 1, 32, 1)]
    a, b = shard.assign_jobs(jobs, 3), shard.assign_jobs(jobs, 3)
    assert a == b and sorted((j.sequence, j.gop) for x in a for j in x) == sorted((j.sequence, j.gop) for j in jobs)
    assert shard.assign_jobs([], 4) == [[], [], [], []]


_GLOO_WORKER = r"""
import os, sys, json
sys.path.insert(0, os.environ["DSVC_ROOT"])
import torch, torch.distributed as dist
from deepsvc_b200 import shard
rank, local_rank, world = shard.init_distributed("gloo")
assert world == 2 and dist.get_backend() == "gloo"
jobs = shard.make_gop_jobs([96] * 7, 32)
mine = shard.assign_jobs(jobs, world)[rank]
# every rank "codes" its jobs (stand-in work), then metrics are merged on the host
frames = sum(j.p_frames for j in mine)
t = 0.010 * frames * (1 + rank)          # fake device time, slower on rank 1
tmax = shard.max_over_ranks(t)
total = shard.sum_over_ranks(frames)
allm = shard.gather_metrics({"rank": rank, "frames": frames})
# training-mode gradient sync: mean over ranks, clamp after the reduction
torch.manual_seed(0)
params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2))]
for p in params:
    p.grad = torch.full_like(p, float(rank + 1) * 2.0)
nb = shard.allreduce_gradients(params, bucket_bytes=16, clamp=1.0)
params2 = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 2))]
for p in params2:
    p.grad = torch.full_like(p, float(rank + 1) * 0.25)
fb = shard.FlatGradBuckets(params2, bucket_bytes=16)
nb2 = fb.allreduce(clamp=1.0)
params2[1].grad.add_(1.0)                      # grads are views of the buckets
# a real model: backward -> zero_grad() (set_to_none=True, Learner.py:177) -> backward -> allreduce.
# The second backward allocates fresh .grad tensors; the buckets must still reduce THEM.
torch.manual_seed(1)
net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Tanh(), torch.nn.Linear(3, 2))
unused = torch.nn.Parameter(torch.ones(2))     # receives no gradient
fb3 = shard.FlatGradBuckets(list(net.parameters()) + [unused], bucket_bytes=32)
opt = torch.optim.SGD(list(net.parameters()) + [unused], lr=0.1)
x = torch.arange(8.0).reshape(2, 4) * (rank + 1)
net(x).sum().backward()
fb3.allreduce()
opt.zero_grad()                                # grads -> None
assert net[0].weight.grad is None
import copy
twin = copy.deepcopy(net)                      # the same step without buckets: this rank's own gradients
twin(x * 0.5).pow(2).sum().backward()
local = [p.grad.clone() for p in twin.parameters()]
net(x * 0.5).pow(2).sum().backward()           # hooks re-bind the fresh grads and start the buckets
fb3.allreduce()
gathered = [None, None]
dist.all_gather_object(gathered, [g.tolist() for g in local])
want = [((torch.tensor(a) + torch.tensor(b)) / 2) for a, b in zip(*gathered)]
ddp_ok = all(torch.allclose(p.grad, w, atol=1e-6) for p, w in zip(net.parameters(), want))
views_ok = all(p.grad.data_ptr() == fb3._views[id(p)].data_ptr() for p in net.parameters())
unused_zero = bool((fb3._views[id(unused)] == 0).all())
out = {"rank": rank, "frames": frames, "tmax": tmax, "total": total, "all": allm,
       "g0": params[0].grad.tolist(), "buckets": nb,
       "f0": params2[0].grad.tolist(), "f1_in_bucket": fb.buckets[1].tolist(), "fbuckets": nb2,
       "ddp_ok": ddp_ok, "views_ok": views_ok, "unused_zero": unused_zero}
if rank == 0:
    print("RESULT " + json.dumps(out), flush=True)
dist.barrier()
dist.destroy_process_group()
"""


def test_world_size_2_gloo(tmp_path):
    """N>1 host path on CPU: job sharding, max-over-ranks timing, metric gather and the
    training-mode gradient all-reduce, over gloo with two processes."""
    import json
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, DSVC_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][0]
    out = json.loads(line[len("RESULT "):])
    assert out["total"] == 21 * 31
    assert out["frames"] == 11 * 31
    assert abs(out["tmax"] - 0.010 * 10 * 31 * 2) < 1e-9       # slowest rank defines the time
    assert [m["frames"] for m in out["all"]] == [11 * 31, 10 * 31]
    assert out["g0"] == [1.0] * 5                              # mean(2,4)=3 -> clamped to 1
    assert out["buckets"] == 2
    assert out["f0"] == [0.375] * 5 and out["fbuckets"] == 2     # mean(0.25, 0.5), in place
    assert out["f1_in_bucket"] == [1.375] * 6                     # p.grad is a view of its bucket
    # backward -> zero_grad(set_to_none) -> backward -> allreduce reduces the step's real gradients
    assert out["ddp_ok"] and out["views_ok"] and out["unused_zero"]
