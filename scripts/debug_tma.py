"""GPU debugging helper: one forced TMA-staged warp on a small shape (run under
compute-sanitizer when chasing a fault)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deepsvc_b200 as d
from deepsvc_b200 import _lib, synthetic
shape = tuple(int(v) for v in (sys.argv[1:5] or (1, 8, 64, 128)))
B, C, H, W = shape
g = torch.Generator().manual_seed(0)
inp = torch.randn(B, C, H, W, generator=g).cuda()
flow = synthetic.smooth_flow(B, H, W, g).cuda()
ref = d.warp_forward(inp, flow, algo=_lib.WARP_GATHER)
torch.cuda.synchronize()
print("gather ok", flush=True)
out = d.warp_forward(inp, flow, algo=_lib.WARP_TMA)
torch.cuda.synchronize()
print("tma ok, max diff", (out - ref).abs().max().item(), "equal", torch.equal(out, ref), flush=True)
