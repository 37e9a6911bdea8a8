#!/usr/bin/env python
"""Small invocations of the newer kernels for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsvc_b200 as d
from deepsvc_b200 import _lib, synthetic
from deepsvc_b200.warp import warp_backward, warp_forward

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for kind in ("smooth", "stress", "border"):
    for (B, C, H, W) in ((1, 16, 48, 128), (2, 8, 33, 68)):
        x = torch.randn(B, C, H, W, generator=g).to(dev)
        f = synthetic.make_flow(kind, B, H, W, g).to(dev)
        go = torch.randn(B, C, H, W, generator=g).to(dev)
        warp_forward(x, f)
        _lib.load().dsvc_set_warp_bwd_algo(_lib.WARP_BWD_STAGED)
        warp_backward(go, x, f, True, True)
        warp_backward(go, x, f, False, True)
        warp_backward(go, x, f, True, False)
        _lib.load().dsvc_set_warp_bwd_algo(_lib.WARP_BWD_AUTO)
im = torch.rand(1, 3, 34, 66, generator=g).to(dev)
d.spynet_level_warp(im, torch.randn(1, 2, 17, 33, generator=g).to(dev))
d.warp_with_mse(im, torch.randn(1, 2, 34, 66, generator=g).to(dev), torch.rand(1, 3, 34, 66, generator=g).to(dev))
d.mc_blend(im, im, im)
torch.cuda.synchronize()
print("sanitize_small: done")
