"""One 1080p (or cfg3) 64-ch warp backward per selected algorithm: the process ncu attaches to.
usage: python scripts/run_bwd_once.py [gather|staged|direct|cell] [cfg3] [stress]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsvc_b200 import _lib, synthetic  # noqa: E402
from deepsvc_b200.warp import warp_backward  # noqa: E402

algo = next((a for a in sys.argv[1:] if a in ("gather", "staged", "direct", "cell")), "gather")
shape = (8, 64, 256, 256) if "cfg3" in sys.argv else (1, 64, 1088, 1920)
kind = "stress" if "stress" in sys.argv else "smooth"
dev = torch.device("cuda:0")
B, C, H, W = shape
g = torch.Generator().manual_seed(1)
inp = torch.randn(B, C, H, W, generator=g).to(dev)
flow = synthetic.make_flow(kind, B, H, W, g).to(dev)
gout = torch.randn(B, C, H, W, generator=g).to(dev)
lib = _lib.load()
lib.dsvc_set_warp_bwd_algo({"direct": 1, "staged": 2, "gather": 3, "cell": 4}[algo])
for _ in range(3):
    warp_backward(gout, inp, flow, True, True)
torch.cuda.synchronize()
print("done", algo, shape, kind)
