"""Multi-GPU partitioning of the hot path.

The path shards by independent units with NO data-path collective: within a GOP the
P-frames are serially dependent (``ref_frame`` / ``feature`` carry,
``test_video.py:368-369``), GOPs are independent (I-frame + ``feature=None`` reset at
``i % GOP == 0``, ``test_video.py:296-297``) and sequences are independent (``:274``).
Unit = (sequence, GOP).  One process per GPU; the only communication is the timing
barrier / max-reduce and a gather of a few floats of metrics.  Training adds a
data-parallel gradient all-reduce (``allreduce_gradients``).
"""
import os
from dataclasses import dataclass
from typing import List, Sequence

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class GopJob:
    sequence: int
    gop: int
    first_frame: int
    n_frames: int  # frames in the GOP, the first one is the I-frame

    @property
    def p_frames(self) -> int:
        return max(self.n_frames - 1, 0)


def make_gop_jobs(frames_per_sequence: Sequence[int], gop: int) -> List[GopJob]:
    """Split every sequence into GOPs (``test_video.py:290-297``: frame i is an I-frame
    iff i % GOP == 0)."""
    jobs = []
    for s, n in enumerate(frames_per_sequence):
        for g, first in enumerate(range(0, n, gop)):
            jobs.append(GopJob(s, g, first, min(gop, n - first)))
    return jobs


def assign_jobs(jobs: Sequence[GopJob], world_size: int) -> List[List[GopJob]]:
    """Static longest-processing-time assignment (cost = P-frames in the GOP);
    deterministic, identical on every rank."""
    order = sorted(range(len(jobs)), key=lambda i: (-jobs[i].p_frames, i))
    loads = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(jobs[i])
        loads[r] += jobs[i].p_frames
    for r in range(world_size):
        out[r].sort(key=lambda j: (j.sequence, j.gop))
    return out


def balance(assignment: Sequence[Sequence[GopJob]]) -> float:
    """mean load / max load: the scaling efficiency bound of a static partition."""
    loads = [sum(j.p_frames for j in a) for a in assignment]
    return (sum(loads) / len(loads)) / max(loads) if max(loads) else 1.0


def dist_env():
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def init_distributed(backend: str = None):
    rank, local_rank, world = dist_env()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU `local_rank`, so that
    pinned host buffers (first touch) and the copy threads live on the GPU's own NUMA node.
    Matters for the host-buffer path at 8 ranks per box; a no-op where NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device-side time of the slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_metrics(local: dict):
    """Per-rank metric dicts on every rank (a few floats; host side)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local)
    return out


def allreduce_gradients(params, bucket_bytes: int = 32 << 20, clamp: float = None):
    """Training-mode data-parallel gradient sync (the only collective on the path):
    mean all-reduce of fp32 grads in flat buckets over NCCL/NVLink, issued
    asynchronously and waited at the end; the reference's element-wise clamp to +-1
    (``Learner.py:1687-1691``) is applied AFTER the reduction so that N GPUs x batch b
    equals one GPU x batch N*b."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    buckets, cur, size = [], [], 0
    for p in params:
        cur.append(p)
        size += p.grad.numel() * p.grad.element_size()
        if size >= bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
    if cur:
        buckets.append(cur)
    pending = []
    for b in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in b])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True) if world > 1 else None
        pending.append((b, flat, work))
    for b, flat, work in pending:
        if work is not None:
            work.wait()
        if world > 1:
            flat.div_(world)
        if clamp is not None:
            flat.clamp_(-clamp, clamp)
        off = 0
        for p in b:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return len(buckets)


class FlatGradBuckets:
    """The same reduction without the per-step gather/scatter copies, overlapped with backward
    ("DDP buckets overlapped with backward"): gradients LIVE in flat fp32 buckets
    (every ``p.grad`` is a view into one), and a post-accumulate-grad hook on every parameter
    starts a bucket's in-place NCCL all-reduce the moment its last gradient of the step has
    been written, while autograd is still producing the earlier layers' gradients.
    ``allreduce()`` after ``backward()`` launches whatever has not been started (parameters
    that received no gradient this step), waits, takes the mean and applies the reference's
    element-wise clamp (``Learner.py:1687-1691``) AFTER the reduction.

    ``optimizer.zero_grad()`` defaults to ``set_to_none=True`` (the reference calls it at
    ``Learner.py:177``), after which autograd allocates fresh ``.grad`` tensors: the hook (and
    ``allreduce()``) copy such a gradient into its bucket view and re-bind ``p.grad`` to the
    view, so the views survive any ``zero_grad`` flavour.  Create it once after the model is
    on its device."""

    def __init__(self, params, bucket_bytes: int = 32 << 20, overlap: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.buckets, self._views, self._bucket_of = [], {}, {}
        self._members = []
        cur, size = [], 0
        for p in self.params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._make(cur)
                cur, size = [], 0
        if cur:
            self._make(cur)
        self._ready = [0] * len(self.buckets)
        self._seen = set()
        self._works = [None] * len(self.buckets)
        self.overlap = overlap
        self._hooks = []
        if overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _make(self, ps):
        flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        off = 0
        bi = len(self.buckets)
        for p in ps:
            n = p.numel()
            view = flat[off:off + n].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            self._views[id(p)] = view
            self._bucket_of[id(p)] = bi
            off += n
        self.buckets.append(flat)
        self._members.append(list(ps))

    def _rebind(self, p) -> None:
        """Make ``p.grad`` the bucket view again (after ``zero_grad(set_to_none=True)`` autograd
        allocated a fresh tensor; after ``p.grad = None`` with no new gradient the view is zeroed)."""
        view = self._views[id(p)]
        g = p.grad
        if g is None:
            view.zero_()
        elif g.data_ptr() != view.data_ptr():
            view.copy_(g)
        else:
            return
        p.grad = view

    def _launch(self, bi: int) -> None:
        if self._works[bi] is not None:
            return
        world = dist.get_world_size() if dist.is_initialized() else 1
        self._works[bi] = (dist.all_reduce(self.buckets[bi], op=dist.ReduceOp.SUM, async_op=True)
                           if world > 1 else True)

    def _on_grad(self, p) -> None:
        if id(p) in self._seen:   # a second accumulation in the same step (shared parameter)
            return
        self._seen.add(id(p))
        self._rebind(p)
        bi = self._bucket_of[id(p)]
        self._ready[bi] += 1
        if self._ready[bi] == len(self._members[bi]):
            self._launch(bi)

    def allreduce(self, clamp: float = None) -> int:
        world = dist.get_world_size() if dist.is_initialized() else 1
        for bi, ps in enumerate(self._members):
            if self._works[bi] is None:
                for p in ps:
                    if id(p) not in self._seen:
                        self._rebind(p)
                self._launch(bi)
        for bi, f in enumerate(self.buckets):
            w = self._works[bi]
            if w is not True:
                w.wait()
            if world > 1:
                f.div_(world)
            if clamp is not None:  # after the reduction (Learner.py:1687-1691 on the reduced gradient)
                f.clamp_(-clamp, clamp)
        self._ready = [0] * len(self.buckets)
        self._works = [None] * len(self.buckets)
        self._seen.clear()
        return len(self.buckets)
