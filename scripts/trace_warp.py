#!/usr/bin/env python
"""Per-CTA timeline of the persistent staged warp kernel (debug build, -DDSVC_TRACE).

Builds a trace-enabled copy of the library under scripts/probe/trace_lib/, runs the 64-ch
1080p feature warp with DSVC_WARP_TRACE set and prints where each role's time goes.
Fields (clock64 ticks): 0/1 scout before/after claim, 2 bbox done, 3 posted, 4/5 issuer
descriptor wait, 6 issuer done, 7 issuer time blocked on empty stages, 8/9 consumer
descriptor wait, 10 consumer prologue done, 11 consumer unit done, 12 unit id,
13 consumer time blocked on full stages, 14 groups."""
import glob
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TL = os.path.join(ROOT, "scripts", "probe", "trace_lib")
LIB = os.path.join(TL, "libdeepsvc_b200.so")


def build():
    os.makedirs(TL, exist_ok=True)
    csrc = os.path.join(ROOT, "deepsvc_b200", "csrc")
    srcs = sorted(glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cpp")))
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
                    "-DDSVC_TRACE", "--threads", "0", "-shared", "-o", LIB] + srcs + ["-lpthread"], check=True)


def main():
    csrc = os.path.join(ROOT, "deepsvc_b200", "csrc")
    stale = not os.path.isfile(LIB) or any(
        os.path.getmtime(f) > os.path.getmtime(LIB) for f in glob.glob(os.path.join(csrc, "*")))
    if stale or "--build" in sys.argv:
        build()
    if "--build-only" in sys.argv:
        return
    import torch
    from deepsvc_b200 import _lib, synthetic
    _lib.LIB_PATH = LIB
    from deepsvc_b200.warp import warp_forward
    dev = torch.device("cuda:0")
    flow_kind = sys.argv[sys.argv.index("--flow") + 1] if "--flow" in sys.argv else "smooth"
    d = synthetic.make_pframe_inputs(B=1, H=1088, W=1920, seed=16)
    if flow_kind == "zero":
        d["flow"].zero_()
    x, f = d["feature"].to(dev), d["flow"].to(dev)
    for _ in range(3):
        warp_forward(x, f)
    torch.cuda.synchronize()
    path = os.path.join(ROOT, "gpurun_out", "warp_trace.bin")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    os.environ["DSVC_WARP_TRACE"] = path
    warp_forward(x, f)
    torch.cuda.synchronize()
    del os.environ["DSVC_WARP_TRACE"]
    analyse(path)


def analyse(path):
    raw = open(path, "rb").read()
    grid, nu, nf, _ = np.frombuffer(raw[:16], dtype=np.int32)
    t = np.frombuffer(raw[16:], dtype=np.int64).reshape(grid, nu, nf).astype(np.float64)
    clk = 1.965e3  # ticks per us (approx., SM clock)
    t0 = t[:, 0, 0].min()
    valid = t[:, :, 11] > 0  # units the consumers finished on the staged fast path
    print(f"grid {grid}; traced fast units {int(valid.sum())}; units/CTA mean {valid.sum(1).mean():.2f}")
    end = np.where(valid, t[:, :, 11], 0).max(1)
    print(f"CTA finish time (us after first claim): min {(end.min()-t0)/clk:.1f} mean {(end.mean()-t0)/clk:.1f} max {(end.max()-t0)/clk:.1f}")
    first = t[:, 0, 9]
    print(f"first descriptor received: mean {(first.mean()-t0)/clk:.2f} us")
    v = valid
    def stat(name, arr):
        a = arr[v] / clk
        print(f"  {name:44s} mean {a.mean():7.2f}  p50 {np.percentile(a,50):7.2f}  p90 {np.percentile(a,90):7.2f}  max {a.max():7.2f}  (us)")
    big = v & (t[:, :, 14] >= 32)
    small = v & (t[:, :, 14] < 32)
    for nm, m in (("whole-tile units", big), ("tail units", small)):
        if not m.any():
            continue
        print(f"{nm}: {int(m.sum())}")
        def st(name, arr):
            a = arr[m] / clk
            print(f"  {name:44s} mean {a.mean():7.2f}  p50 {np.percentile(a,50):7.2f}  p90 {np.percentile(a,90):7.2f}  max {a.max():7.2f}  (us)")
        st("scout: claim (atomic)", t[:, :, 1] - t[:, :, 0])
        st("scout: bbox", t[:, :, 2] - t[:, :, 1])
        st("scout: post (waits for a free descriptor)", t[:, :, 3] - t[:, :, 2])
        st("issuer: waits for descriptor", t[:, :, 5] - t[:, :, 4])
        st("issuer: unit issue span", t[:, :, 6] - t[:, :, 5])
        st("issuer:   of which blocked on empty", t[:, :, 7])
        st("consumer: waits for descriptor", t[:, :, 9] - t[:, :, 8])
        st("consumer: prologue (flow, taps)", t[:, :, 10] - t[:, :, 9])
        st("consumer: channel loop", t[:, :, 11] - t[:, :, 10])
        st("consumer:   of which blocked on full", t[:, :, 13])
        st("consumer: loop time per group", (t[:, :, 11] - t[:, :, 10]) / np.maximum(t[:, :, 14], 1))
    # one CTA's timeline
    c = 0
    print("CTA 0 timeline (us): unit id | desc recv | prologue done | loop done | blocked full | groups")
    for u in range(nu):
        if t[c, u, 9] == 0:
            break
        print(f"  {int(t[c,u,12]):5d} {(t[c,u,9]-t0)/clk:8.2f} {(t[c,u,10]-t0)/clk:8.2f} {(t[c,u,11]-t0)/clk:8.2f} {t[c,u,13]/clk:8.2f} {int(t[c,u,14]):3d}")


if __name__ == "__main__":
    main()
