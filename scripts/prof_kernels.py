#!/usr/bin/env python
"""Per-kernel timing of the hot path's launches at one frame size (CUDA events, warm,
inputs larger than L2 or L2 flushed between launches).  Also the command ncu wraps:

    python scripts/prof_kernels.py --what feature --iters 20
    ncu --set full --clock-control none --import-source on -k regex:warp_fwd_tma -s 3 -c 1 \
        -o gpurun_out/prof python scripts/prof_kernels.py --what feature --iters 4

--what: feature | frame | spynet | gc | eb | bwd_feature | bwd_frame | all
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from deepsvc_b200 import _lib, synthetic  # noqa: E402
from deepsvc_b200.warp import warp_forward, warp_backward  # noqa: E402


def timeit(fn, iters, flush):
    evs = []
    for _ in range(3):
        fn()
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    t = [a.elapsed_time(b) for a, b in evs]
    return statistics.median(t), min(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--height", type=int, default=1088)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--flow", default="smooth")  # smooth | stress | border | zero | shift | nojitter
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--no-flush", action="store_true")
    a = ap.parse_args()
    _lib.load()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    B, H, W = a.batch, a.height, a.width
    special = a.flow in ("zero", "shift", "nojitter")
    cpu_in = synthetic.make_pframe_inputs(B=B, H=H, W=W, seed=16, flow_kind="smooth" if special else a.flow)
    if a.flow == "zero":
        cpu_in["flow"].zero_()
    elif a.flow == "shift":
        cpu_in["flow"][:, 0] = 5.3
        cpu_in["flow"][:, 1] = -3.6
    elif a.flow == "nojitter":
        cpu_in["flow"] = synthetic.smooth_flow(B, H, W, torch.Generator().manual_seed(3), jitter=0.0)
    d = synthetic.to_device(cpu_in, dev)
    flush = None if a.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    algo = {"auto": _lib.WARP_AUTO, "gather": _lib.WARP_GATHER, "tma": _lib.WARP_TMA}[a.algo]
    res = {}

    def rec(name, ms, nbytes):
        med, mn = ms
        res[name] = {"ms_median": med, "ms_min": mn, "alg_bytes": nbytes,
                     "gbs_median": nbytes / med / 1e6, "gbs_best": nbytes / mn / 1e6}
        print(f"{name:28s} median {med*1e3:9.1f} us  best {mn*1e3:9.1f} us  "
              f"{nbytes/1e6:9.1f} MB  {nbytes/med/1e6:8.1f} GB/s (best {nbytes/mn/1e6:8.1f})", flush=True)

    want = lambda k: a.what in ("all", k)
    if want("feature"):
        x, f = d["feature"], d["flow"]
        rec("warp_fwd C=64", timeit(lambda: warp_forward(x, f, algo=algo), a.iters, flush),
            4 * B * H * W * (2 * 64 + 2))
    if want("frame"):
        x, f = d["ref_frame"], d["flow"]
        rec("warp_fwd C=3", timeit(lambda: warp_forward(x, f, algo=algo), a.iters, flush), 4 * B * H * W * 8)
    if want("spynet"):
        for x, f in zip(d["pyr_img"], d["pyr_flow"]):
            rec(f"warp_fwd C=3 {x.shape[2]}x{x.shape[3]}", timeit(lambda: warp_forward(x, f, algo=algo), a.iters, flush),
                4 * B * x.shape[2] * x.shape[3] * 8)
    if want("bwd_feature"):
        x, f = d["feature"], d["flow"]
        g = torch.randn_like(x)
        rec("warp_bwd C=64 both", timeit(lambda: warp_backward(g, x, f, True, True), a.iters, flush),
            4 * B * H * W * (3 * 64 + 4))
        rec("warp_bwd C=64 flow-only", timeit(lambda: warp_backward(g, x, f, False, True), a.iters, flush),
            4 * B * H * W * (2 * 64 + 4))
        rec("warp_bwd C=64 input-only", timeit(lambda: warp_backward(g, x, f, True, False), a.iters, flush),
            4 * B * H * W * (2 * 64 + 2))
    if want("bwd_frame"):
        x, f = d["ref_frame"], d["flow"]
        g = torch.randn_like(x)
        rec("warp_bwd C=3 flow-only", timeit(lambda: warp_backward(g, x, f, False, True), a.iters, flush),
            4 * B * H * W * (2 * 3 + 4))
    if want("fused"):
        import torch.nn.functional as F
        import deepsvc_b200 as dsvc
        with torch.no_grad():
            for k in (1, 2, 3):  # SpyNet levels 1..3 warp with the upsampled flow of the level below
                img = d["pyr_img"][k]
                h, w = img.shape[2], img.shape[3]
                coarse = d["pyr_flow"][k][:, :, ::2, ::2].contiguous()
                def unfused():
                    fu = F.interpolate(coarse, (h, w), mode="bilinear", align_corners=False) * 2.0
                    return fu, warp_forward(img, fu)
                nb = 4 * B * h * w * (2 * 3 + 2) + 4 * B * h * w * 2 // 4
                rec(f"spynet level {h}x{w} unfused", timeit(unfused, a.iters, flush), nb)
                rec(f"spynet level {h}x{w} fused", timeit(lambda: dsvc.spynet_level_warp(img, coarse), a.iters, flush), nb)
            ref, fl, cur = d["ref_frame"], d["flow"], torch.rand_like(d["ref_frame"])
            def unfused2():
                wv = warp_forward(ref, fl)
                return wv, torch.mean((wv - cur).pow(2))
            nb = 4 * B * H * W * (3 * 3 + 2)
            rec("frame warp + mse unfused", timeit(unfused2, a.iters, flush), nb)
            rec("frame warp + mse fused", timeit(lambda: dsvc.warp_with_mse(ref, fl, cur), a.iters, flush), nb)
    if want("gc") or want("eb"):
        import deepsvc_b200 as dsvc
        gc = dsvc.GaussianConditional(None).to(dev).eval()
        eb = dsvc.EntropyBottleneck(96).to(dev).eval()
        y, s, m, z = d["res_y"], d["res_scales"], d["res_means"], d["res_z"]
        with torch.no_grad():
            if want("gc"):
                ys, ss, ms_ = y.chunk(8, 1)[0], s.chunk(8, 1)[0], m.chunk(8, 1)[0]
                rec("gc_fwd slice 12ch", timeit(lambda: gc(ys, ss, ms_), a.iters, None), 20 * ys.numel())
                rec("gc_fwd all 96ch", timeit(lambda: gc(y, s, m), a.iters, None), 20 * y.numel())
            if want("eb"):
                rec("eb_fwd 96ch", timeit(lambda: eb(z), a.iters, None), 12 * z.numel())
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"prof_kernels_{a.what}_{a.algo}.json"), "w") as fh:
        json.dump({"args": vars(a), "results": res}, fh, indent=1)


if __name__ == "__main__":
    main()
