"""GPU check + timing of the warp backward kernels (direct / staged / gather) against each other.
usage: python scripts/check_bwd.py [--small] [--time]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepsvc_b200 import _lib, synthetic  # noqa: E402
from deepsvc_b200.warp import warp_backward  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
ALGOS = {"auto": _lib.WARP_BWD_AUTO, "direct": _lib.WARP_BWD_DIRECT, "staged": _lib.WARP_BWD_STAGED,
         "gather": _lib.WARP_BWD_GATHER, "cell": _lib.WARP_BWD_CELL}


CHECKED = [a for a in ("gather", "cell") if ("--only" not in sys.argv or a in sys.argv)]


def run(algo, gout, inp, flow, gi=True, gf=True):
    _lib.check(lib.dsvc_set_warp_bwd_algo(ALGOS[algo]), "algo")
    try:
        return warp_backward(gout, inp, flow, gi, gf)
    finally:
        lib.dsvc_set_warp_bwd_algo(0)


def check(shape, kind, seed=0):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(seed + H + W)
    inp = torch.randn(B, C, H, W, generator=g).to(dev)
    flow = synthetic.make_flow(kind, B, H, W, g).to(dev)
    gout = torch.randn(B, C, H, W, generator=g).to(dev)
    ref = run("direct", gout, inp, flow)
    auto = run("auto", gout, inp, flow)
    torch.cuda.synchronize()
    for a, b in zip(auto, ref):
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item()), "auto (scout + staged | direct)"
    worst = 0.0
    for algo in CHECKED:
        got = run(algo, gout, inp, flow)
        out = []
        for a, b in zip(got, ref):
            out.append((a - b).abs().max().item() / max(1.0, b.abs().max().item()))
        got2 = run(algo, gout, inp, flow, True, False)
        out.append((got2[0] - ref[0]).abs().max().item() / max(1.0, ref[0].abs().max().item()))
        print(f"{shape} {kind} {algo}: rel err gin {out[0]:.2e} gflow {out[1]:.2e} gin-only {out[2]:.2e}", flush=True)
        worst = max(worst, max(out))
    return worst


def timeit(shape, kind, algos=("auto", "direct", "staged", "gather", "cell"), n=10):
    B, C, H, W = shape
    g = torch.Generator().manual_seed(1)
    inp = torch.randn(B, C, H, W, generator=g).to(dev)
    flow = synthetic.make_flow(kind, B, H, W, g).to(dev)
    gout = torch.randn(B, C, H, W, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nbytes = 4 * B * H * W * (3 * C + 4)
    for a in algos:
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(a, gout, inp, flow)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        print(f"time {shape} {kind} {a}: {ms * 1e3:.1f} us (incl. allocation / zero-fill)  {nbytes / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    EXIT = 0
    worst = 0.0
    small = [((1, 8, 20, 40), "smooth"), ((1, 8, 32, 64), "smooth"), ((2, 16, 40, 64), "smooth"), ((1, 9, 33, 100), "border"), ((1, 8, 16, 68), "stress")]
    full = small + [((1, 64, 128, 192), k) for k in ("smooth", "stress", "border")] + \
        [((8, 64, 64, 64), "smooth"), ((2, 16, 272, 480), "smooth"), ((1, 64, 1088, 1920), "smooth")]
    for shape, kind in (small if "--small" in sys.argv else full):
        worst = max(worst, check(shape, kind))
    print("worst rel err", worst)
    if "--time" in sys.argv:
        algos = tuple(a for a in ALGOS if "--only" not in sys.argv or a in sys.argv)
        timeit((1, 64, 1088, 1920), "smooth", algos)
        timeit((8, 64, 256, 256), "smooth", algos)
        timeit((1, 64, 1088, 1920), "stress", algos)
        if "cell" in algos:
            timeit((1, 64, 1088, 1920), "border", algos)
            timeit((1, 64, 1088, 1920), "gentle", algos)
    EXIT = 0 if worst <= 1e-4 else 1


def time_parts(shape=(1, 64, 1088, 1920), kind="smooth", n=10):
    """gin-only and gflow-only launches per algorithm (is a two-kernel split worth it?)."""
    B, C, H, W = shape
    g = torch.Generator().manual_seed(1)
    inp = torch.randn(B, C, H, W, generator=g).to(dev)
    flow = synthetic.make_flow(kind, B, H, W, g).to(dev)
    gout = torch.randn(B, C, H, W, generator=g).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for a in ("direct", "staged", "gather"):
        for gi, gf in ((True, False), (False, True)):
            ts = []
            for _ in range(n):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(a, gout, inp, flow, gi, gf)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print(f"parts {shape} {a} gin={gi} gflow={gf}: {ts[len(ts) // 2] * 1e3:.1f} us", flush=True)


if __name__ == "__main__" and "--parts" in sys.argv:
    time_parts()
    time_parts((8, 64, 256, 256))

if __name__ == "__main__":
    sys.exit(EXIT)
