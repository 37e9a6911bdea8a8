#!/usr/bin/env python
"""Does a short launch (3-ch frame warp) become resident next to the persistent 64-ch feature warp?
Two streams, CUDA events: reports each kernel's own duration and the pair's makespan."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deepsvc_b200 import _lib, synthetic
from deepsvc_b200.hotpath import PFrameHotPath
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
_lib.load()
cpu_in = synthetic.make_pframe_inputs(B=1, H=1088, W=1920, seed=16)
hp = PFrameHotPath(synthetic.to_device(cpu_in, dev), bench.build_models(dev))
calls = {name: (fn, a) for fn, a, name in hp._calls}
feat = [c for c in hp._calls if c[2].startswith("warp_c64")][0]
frame = [c for c in hp._calls if c[2] == "warp_c3_1088x1920"][-1]
gc = [c for c in hp._calls if c[2] == "gc_res"][0]
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
hp.run(); torch.cuda.synchronize()

def go(second, order):
    res = []
    for it in range(12):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record(torch.cuda.current_stream())
        s1.wait_event(e[0]); s2.wait_event(e[0])
        def a():
            e[1].record(s1); feat[0](*feat[1], s1.cuda_stream); e[2].record(s1)
        def b():
            e[3].record(s2)
            for _ in range(second[1]):
                second[0][0](*second[0][1], s2.cuda_stream)
            e[4].record(s2)
        (a(), b()) if order == "feat_first" else (b(), a())
        torch.cuda.synchronize()
        res.append((e[1].elapsed_time(e[2]), e[3].elapsed_time(e[4]), max(e[0].elapsed_time(e[2]), e[0].elapsed_time(e[4]))))
    res = res[2:]
    return [statistics.median(r[i] for r in res) * 1e3 for i in range(3)]

for name, second in (("frame warp x1", (frame, 1)), ("frame warp x4", (frame, 4)), ("gc x8", (gc, 8))):
    for order in ("feat_first", "short_first"):
        f, s, tot = go(second, order)
        print(f"{name:14s} {order:12s} feature {f:7.1f} us  short {s:7.1f} us  makespan {tot:7.1f} us", flush=True)
