"""Host time per eager drop-in call (no synchronisation inside the loop): what an unmodified
DeepSVC.forward pays per hot-path call on top of the kernel itself.  r01: gc 57-60 us, eb 69 us."""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deepsvc_b200 as d
from deepsvc_b200 import _lib, synthetic

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
y, s, m = (t.to(dev) for t in synthetic.make_latents(1, 12, 68, 120, g))
z = (torch.randn(1, 96, 17, 30, generator=g) * 3).to(dev)
x3 = torch.rand(1, 3, 272, 480, generator=g).to(dev)
f3 = synthetic.smooth_flow(1, 272, 480, g).to(dev)
gc = d.GaussianConditional(None).to(dev).eval()
eb = d.EntropyBottleneck(96).to(dev).eval()


def host_us(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt / n * 1e6


res = {"torch_ops_layer": _lib.torch_ops() is not None}
with torch.no_grad():
    res["gc(y, scale, mu) [1,12,68,120]"] = host_us(lambda: gc(y, s, m))
    res["eb(z) [1,96,17,30]"] = host_us(lambda: eb(z))
    res["torch_warp 3ch 272x480"] = host_us(lambda: d.torch_warp(x3, f3))
    res["gc.quantize_and_index"] = host_us(lambda: gc.quantize_and_index(y, s, m)) if gc.scale_table.numel() else None
    res["stock torch: y - m (one elementwise op, for scale)"] = host_us(lambda: y - m)
print(json.dumps(res, indent=1))
