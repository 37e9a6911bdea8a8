#!/usr/bin/env python
"""Frame time of the hot-path DAG with subsets of its launches (CUDA graph replays, CUDA events):
which launches add to the 64-ch feature warp's time and which hide under it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deepsvc_b200 import _lib, synthetic
from deepsvc_b200.hotpath import PFrameHotPath

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
cpu_in = synthetic.make_pframe_inputs(B=1, H=1088, W=1920, seed=16)
gpu_in = synthetic.to_device(cpu_in, dev)
models = bench.build_models(dev)


FUSE = "--fuse" in sys.argv


def timed(keep, label):
    hp = PFrameHotPath(gpu_in, models, fuse_frame_warp=FUSE)
    hp._calls = [c for c in hp._calls if keep(c[2]) or c[2] == "bits_finalize"]
    hp.capture()
    for _ in range(20):
        hp.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(300):
        hp.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label:46s} {len(hp._calls):3d} launches  {e0.elapsed_time(e1) / 300 * 1e3:8.1f} us/frame", flush=True)


feat = lambda n: n.startswith("warp_c64")
full3 = lambda n: n == "warp_c3_1088x1920"
pyr = lambda n: n.startswith("warp_c3") and not full3(n)
ent = lambda n: n.startswith("gc_") or n.startswith("eb_")
timed(feat, "feature warp only")
timed(lambda n: feat(n) or full3(n), "feature + two full-res 3-ch warps")
timed(lambda n: feat(n) or full3(n) or pyr(n), "all six warps")
timed(lambda n: feat(n) or ent(n), "feature + 18 entropy launches")
timed(lambda n: full3(n) or pyr(n), "five 3-ch warps only")
timed(ent, "18 entropy launches only")
timed(lambda n: True, "whole frame")
timed(lambda n: feat(n) or (full3(n) and True) or ent(n), "feature + full-res 3-ch + entropy")
timed(lambda n: feat(n) or pyr(n), "feature + the three small pyramid warps")
