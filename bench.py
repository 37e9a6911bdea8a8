#!/usr/bin/env python
"""Benchmark of the DeepSVC warp + entropy P-frame hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one 1088x1920 P-frame (B=1, BASELINE.json configs[1]) through the hot
path of ``DeepSVC.forward``: 4 SpyNet warps + frame warp + 64-ch feature warp + 16
GaussianConditional slice calls + 2 EntropyBottleneck calls + bit sums (25 launches;
conv transforms excluded, their outputs are pre-generated synthetic tensors).

Prints ONE JSON line (rank 0).  Keys beyond the driver's base contract:
  roofline      the 64-ch feature warp: algorithmic bytes / CUDA-event time vs the measured
                HBM copy peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (restatement of the reference's CPU path) timed on this
                box's host cores on a bounded sample (rank 0, N=1 only)
  e2e           the same metric through the host-buffer API (pinned host inputs,
                H2D + kernels + D2H inside the timed region)
  clocks        NVML samples taken during the timed regions
``--impl reference`` times the reference's own CPU implementation of the path (oracle
port: torch CPU ops op-for-op as modules.py / compressai issue them), all host threads.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "1080p P-frames/sec (warp+entropy path)"
UNIT = "frames/s"
H, W, B = 1088, 1920, 1
MIN_TIMED_S = 0.5   # the K-step block is repeated until the timed region is at least this long
# frame sizes BASELINE.json's configs name (padded to multiples of 64, modules.py:76-89)
SIZES = {"cfg1": (256, 448), "cfg2": (1088, 1920), "cfg5": (2176, 3840)}


def labels(Hh, Ww, flow="smooth"):
    """(metric, workload, size tag) for a frame size: the line is labelled from the size it ran
    at, never from a constant.  Sizes outside BASELINE.json's configs are refused."""
    tag = {v: k for k, v in SIZES.items()}.get((Hh, Ww))
    if tag is None:
        raise SystemExit(f"bench.py: {Ww}x{Hh} is not a BASELINE.json frame size "
                         f"(use one of {sorted(SIZES.values())})")
    name = {"cfg1": "448x256 (Vimeo90k-shaped)", "cfg2": "1920x1088 (padded 1080p)",
            "cfg5": "3840x2176 (padded 2160p)"}[tag]
    metric = {"cfg1": "448x256 P-frames/sec (warp+entropy path)", "cfg2": METRIC,
              "cfg5": "2160p P-frames/sec (warp+entropy path)"}[tag]
    h16, w16, h64, w64 = Hh // 16, Ww // 16, Hh // 64, Ww // 64
    workload = (f"{tag}: {name} P-frame, B=1, DeepSVC.forward hot path = 4 SpyNet "
                "3-ch warps + 3-ch frame warp + 64-ch feature warp + 16 GaussianConditional slices "
                f"(8x8ch mv, 8x12ch res @{h16}x{w16}) + 2 EntropyBottleneck (64/96ch @{h64}x{w64}) + bit sums; "
                f"{flow} flow" + (" (SpyNet-like)" if flow == "smooth" else ""))
    return metric, workload, tag


WORKLOAD = labels(H, W)[1]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--flow", default="smooth", choices=["smooth", "stress", "border", "gentle"],
                    help="smooth (default, SURVEY 8d primary), stress, border; gentle = a low-gradient flow for one secondary line")
    ap.add_argument("--algo", default="auto", choices=["auto", "gather", "tma"])
    ap.add_argument("--branches4", action="store_true",
                    help="capture the 4-branch DAG (feature | 3-ch warps | mv | res entropy) instead "
                         "of one graph branch per launch")
    ap.add_argument("--serial", action="store_true",
                    help="capture the frame's 25 launches in serial order instead of as a DAG")
    ap.add_argument("--no-fuse-frame-warp", dest="fuse_frame_warp", action="store_false",
                    help="default: the 3-ch frame warp (video_model.py:37) rides on the 64-ch feature warp's launch "
                         "(modules.py:429: same flow, bit-identical outputs, 24 launches per frame); this flag "
                         "times the 25-launch sequence of separate calls instead (also reported in config either way)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: min(steps, 40)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "codec", "dropin"],
                    help="cfg2 (default, the headline line): 1080p P-frame forward path; cfg1 / cfg5: the same "
                         "path at 448x256 (rotated input sets, see config.l2) / 3840x2176; cfg3: training "
                         "frame-step (B=8, 256x256, forward + backward); cfg4: GOP sharding; codec: symbol "
                         "pipeline + range coder; dropin: the unmodified reference DeepSVC.forward stock vs "
                         "patched (needs oracle/_ref) -- all secondary lines")
    ap.add_argument("--threads", type=int, default=0, help="--workload codec: host coder threads (default: all cores)")
    ap.add_argument("--train", action="store_true", help="--workload dropin: time a training step instead of inference")
    ap.add_argument("--sets", type=int, default=0,
                    help="distinct input sets rotated between steps (default: 1, or 8 at cfg1 whose 69 MB "
                         "working set would otherwise sit in the 126 MB L2)")
    ap.add_argument("--gop", type=int, default=32, help="cfg4: GOP size (32 = BASELINE, 12 = test_video.py:22)")
    ap.add_argument("--height", type=int, default=H)
    ap.add_argument("--width", type=int, default=W)
    return ap.parse_args()


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls NVML (SM clock, throttle reasons) in a thread while a timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap",
               0x8: "hw_slowdown", 0x10: "sync_boost", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index=0, period=0.002):
        self.samples, self.reason_bits = [], 0
        self.period, self._stop, self._thr = period, threading.Event(), None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._poll, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        reasons = [n for b, n in self.REASONS.items() if self.reason_bits & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ----------------------------------------------------------------------------- models
def build_models(device, seed=0):
    """Random-init entropy models of the reference's shapes (image_model.py:148-149 with
    N=64 / N=96), tanh gates and medians perturbed so that every term is exercised."""
    import torch
    import deepsvc_b200 as dsvc
    g = torch.Generator().manual_seed(seed)
    models = {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb = dsvc.EntropyBottleneck(ch)
        with torch.no_grad():
            for i in range(4):
                f = getattr(eb, f"_factor{i}")
                f.copy_(torch.randn(f.shape, generator=g) * 0.1)
            eb.quantiles[:, 0, 1] = torch.randn(ch, generator=g) * 0.3
        models[name] = (eb.to(device).eval(), dsvc.GaussianConditional(None).to(device).eval())
    return models


def oracle_models(models):
    """CPU oracle twins of the GPU models (same parameters) for the cpu_baseline legs."""
    from oracle import reference_ops as R
    out = {}
    for name, (eb, _) in models.items():
        eb_o = R.EntropyBottleneck(eb.channels)
        eb_o.load_state_dict({k: v.cpu() for k, v in eb.state_dict().items()}, strict=False)
        out[name] = (eb_o.eval(), R.GaussianConditional(None).eval())
    return out


# ----------------------------------------------------------------------------- CPU arm
def crop_rows(inputs, frac_rows):
    """Bounded sample: the top `frac_rows` (multiple of 64) rows of every tensor."""
    import torch
    out = {}
    for k, v in inputs.items():
        if isinstance(v, list):
            out[k] = [t[:, :, : max(1, int(t.shape[2] * frac_rows / inputs["ref_frame"].shape[2]))].contiguous() for t in v]
        else:
            out[k] = v[:, :, : max(1, int(v.shape[2] * frac_rows / inputs["ref_frame"].shape[2]))].contiguous()
    return out


def time_cpu_path(cpu_inputs, models_o, steps, warmup, threads, budget_s):
    """Times oracle.pframe_hotpath on the host. Returns (frames/s, description, n_steps).
    `cpu_inputs`: one input set, or a list of sets used in rotation (the `config.l2` policy)."""
    import torch
    from oracle import reference_ops as R
    torch.set_num_threads(threads)
    sets = cpu_inputs if isinstance(cpu_inputs, list) else [cpu_inputs]
    cpu_inputs = sets[0]
    Hh = cpu_inputs["ref_frame"].shape[2]
    with torch.no_grad():
        t0 = time.perf_counter()
        probe_rows = 64
        R.pframe_hotpath(crop_rows(cpu_inputs, probe_rows), models_o)
        t_probe = time.perf_counter() - t0
    est_full = t_probe * Hh / probe_rows
    total = max(steps + warmup, 1)
    rows = Hh
    if est_full * total > budget_s:
        rows = int(budget_s / (est_full * total) * Hh) // 64 * 64
        rows = min(max(rows, 64), Hh)
    samples = [crop_rows(ci, rows) if rows < Hh else ci for ci in sets]
    frac = rows / Hh
    with torch.no_grad():
        for i in range(warmup):
            R.pframe_hotpath(samples[i % len(samples)], models_o)
        t0 = time.perf_counter()
        for i in range(steps):
            R.pframe_hotpath(samples[i % len(samples)], models_o)
        dt = time.perf_counter() - t0
    fps = steps * frac / dt
    desc = (f"{steps} steps (+{warmup} warm-up) of the top {rows}/{Hh} rows of every tensor of one "
            f"{cpu_inputs['ref_frame'].shape[3]}x{Hh} P-frame (value scaled by {frac:.4f}), torch {threads} threads, {dt:.1f}s")
    return fps, desc, dt / steps


def l2_note(total_bytes, nsets):
    """How the timed region keeps its reads out of L2 (part of `config`: a property of the workload)."""
    if nsets == 1:
        return (f"inputs larger than L2: {total_bytes / 1e6:.0f} MB working set per frame vs 126 MB L2, "
                "no flush needed")
    return (f"{nsets} distinct input/output sets replayed in rotation ({nsets} x {total_bytes / 1e6:.0f} MB "
            "> 126 MB L2): every replay reads its frame from HBM")


def default_sets(total_bytes):
    return 8 if total_bytes < (256 << 20) else 1


def frame_size(args):
    if args.workload in SIZES:
        if (args.height, args.width) != (H, W) and (args.height, args.width) != SIZES[args.workload]:
            raise SystemExit("bench.py: --height/--width contradict --workload")
        return SIZES[args.workload] if args.workload != "cfg2" else (args.height, args.width)
    return args.height, args.width


def run_reference(args):
    import torch
    from deepsvc_b200 import synthetic
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    Hh, Ww = frame_size(args)
    metric, workload, _ = labels(Hh, Ww, args.flow)
    nsets = args.sets or default_sets(synthetic.pframe_algorithmic_bytes(B, Hh, Ww)["total"])
    cpu_in = [synthetic.make_pframe_inputs(B=B, H=Hh, W=Ww, seed=16 + 100 * i, flow_kind=args.flow) for i in range(nsets)]
    models = build_models("cpu")
    models_o = oracle_models(models)
    fps, desc, s_per_step = time_cpu_path(cpu_in, models_o, args.steps, max(args.warmup, 1) if args.steps > 3 else 1,
                                          threads, budget_s=150.0)
    line = {
        "impl": "reference", "metric": metric, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload, "flow": args.flow,
                   "l2": l2_note(synthetic.pframe_algorithmic_bytes(B, Hh, Ww)["total"], nsets)},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "host": {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def timed_blocks(replay, steps, barrier, dev, min_s=MIN_TIMED_S, max_blocks=400):
    """Times blocks of exactly `steps` calls of `replay(i)` with CUDA events on the current
    stream until the timed region is at least `min_s` long; returns (median block ms, per-block
    ms list).  Bracketed by barrier + synchronize on both sides."""
    import torch
    barrier()
    blocks, total, i = [], 0.0, 0
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            replay(i)
            i += 1
        e1.record()
        e1.synchronize()
        blocks.append(e0.elapsed_time(e1))
        total += blocks[-1]
        if total >= min_s * 1e3 or len(blocks) >= max_blocks:
            break
    barrier()
    return statistics.median(blocks), blocks


def time_gpu_eager(fn, dev, steps=10, warmup=3):
    """ms per call of an eager GPU function (CUDA events around `steps` calls)."""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / steps


def run_ours(args):
    import torch
    import deepsvc_b200  # noqa: F401
    from deepsvc_b200 import _lib, shard, synthetic
    from deepsvc_b200.hotpath import HostSession, PFrameHotPath, pframe_eager

    rank, local_rank, world = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for --impl ours)")
    _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_bound = shard.bind_to_gpu_numa_node(local_rank) if world > 1 else False
    Hh, Ww = frame_size(args)
    metric, workload, tag = labels(Hh, Ww, args.flow)
    algo = {"auto": _lib.WARP_AUTO, "gather": _lib.WARP_GATHER, "tma": _lib.WARP_TMA}[args.algo]
    bytes_alg = synthetic.pframe_algorithmic_bytes(B, Hh, Ww)
    # L2: 126 MB.  A frame whose working set fits is run over `nsets` distinct input/output sets
    # in rotation so that every replay finds its data in HBM, not in L2
    nsets = args.sets or default_sets(bytes_alg["total"])

    models = build_models(dev)
    cpu_sets = [synthetic.make_pframe_inputs(B=B, H=Hh, W=Ww, seed=16 + rank + 100 * i, flow_kind=args.flow)
                for i in range(nsets)]
    cpu_in = cpu_sets[0]
    hps = []
    for ci in cpu_sets:
        hp = PFrameHotPath(synthetic.to_device(ci, dev), models, warp_algo=algo, fuse_frame_warp=args.fuse_frame_warp)
        hp.capture(dag=False)                                   # "serial": DeepSVC.forward's own order
        hp.capture(dag=True if args.branches4 else "wide")      # data-dependency DAG
        hps.append(hp)
    # the other launch sequence (frame warp fused / not fused), first input set only, for config
    hp_alt = PFrameHotPath(hps[0].inputs, models, warp_algo=algo, fuse_frame_warp=not args.fuse_frame_warp)
    hp_alt.capture(dag=False)
    hp_alt.capture(dag=True if args.branches4 else "wide")
    hp, gpu_in = hps[0], hps[0].inputs
    main_mode = "serial" if args.serial else ("branches4" if args.branches4 else "wide")

    def barrier():
        if world > 1:
            torch.distributed.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    # ---- main timed region: blocks of K graph replays, inputs resident in HBM
    for i in range(max(args.warmup, 3)):
        hps[i % nsets].replay(main_mode)
    sampler = ClockSampler(local_rank)
    with sampler:
        ms_local, blocks = timed_blocks(lambda i: hps[i % nsets].replay(main_mode), args.steps, barrier, dev)
    ms = shard.max_over_ranks(ms_local, dev)
    value = world * args.steps / (ms * 1e-3)
    bpp = hp.results()["bpp"]
    # the same frame in the serial order of DeepSVC.forward (what a caller that interleaves the
    # conv transforms can issue); reported next to the DAG number
    other = "serial" if main_mode != "serial" else "wide"
    for i in range(3):
        hps[i % nsets].replay(other)
    with sampler:
        ms_other, _ = timed_blocks(lambda i: hps[i % nsets].replay(other), args.steps, barrier, dev)
    ms_other = shard.max_over_ranks(ms_other, dev)
    order_ms = {main_mode: ms / args.steps, other: ms_other / args.steps}
    alt_ms = {}
    for mode in ("serial", "branches4" if args.branches4 else "wide"):
        for _ in range(3):
            hp_alt.replay(mode)
        with sampler:
            m_, _ = timed_blocks(lambda i: hp_alt.replay(mode), args.steps, barrier, dev, min_s=0.2)
        alt_ms["serial" if mode == "serial" else "dag"] = shard.max_over_ranks(m_, dev) / args.steps

    # ---- dominant kernel (64-ch feature warp) timed live with CUDA events, same stream,
    #      inside K eager steps of the whole frame (so caches/clocks see the full step)
    feat_idx = [i for i, c in enumerate(hp._calls) if c[2].startswith("warp_c64")][0]
    st = torch.cuda.current_stream(dev)
    n_k = 200   # eager frames for the kernel timing (independent of --steps: 20 frames gave a 4 % low reading)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_k)]
    for i in range(n_k):
        for j, (fn, a, name) in enumerate(hps[i % nsets]._calls):
            if j == feat_idx:
                evs[i][0].record(st)
            err = fn(*a, st.cuda_stream)
            if err:
                _lib.check(err, name)
            if j == feat_idx:
                evs[i][1].record(st)
    torch.cuda.synchronize(dev)
    k_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in evs[1:] or evs)
    peak, peak_src = measured_peak()
    # (with the frame warp riding on the same launch the kernel moves the frame's 3 planes too)
    k_bytes = bytes_alg["feature"] + (bytes_alg["frame"] if args.fuse_frame_warp else 0)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    frame_gbs = bytes_alg["total"] / (ms_local / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": f"warp_fwd_persist (64-ch feature warp" +
                (" + the 3-ch frame warp on the same flow" if args.fuse_frame_warp else "") + f", {Hh}x{Ww})",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": None,
                "algorithmic_bytes_per_launch": k_bytes, "kernel_ms": k_ms,
                "frac_of_nominal_8tbs": achieved / 8000.0,
                "whole_frame": {"algorithmic_bytes": bytes_alg["total"], "achieved_gbs": frame_gbs,
                                "frac_of_peak": frame_gbs / peak, "frac_of_nominal_8tbs": frame_gbs / 8000.0,
                                "serial_order_frac_of_peak":
                                    bytes_alg["total"] / (order_ms["serial"] * 1e-3) / 1e9 / peak}}
    # the other launch sequence's 64-ch warp, timed the same way (continuity with round 1, whose
    # default line timed the separate 64-ch launch)
    alt_idx = [i for i, c in enumerate(hp_alt._calls) if c[2].startswith("warp_c64")][0]
    n_a = min(n_k, 50)
    evs_a = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_a)]
    for i in range(n_a):
        for j, (fn, a, name) in enumerate(hp_alt._calls):
            if j == alt_idx:
                evs_a[i][0].record(st)
            err = fn(*a, st.cuda_stream)
            if err:
                _lib.check(err, name)
            if j == alt_idx:
                evs_a[i][1].record(st)
    torch.cuda.synchronize(dev)
    a_ms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in evs_a[1:] or evs_a)
    a_bytes = bytes_alg["feature"] + (0 if args.fuse_frame_warp else bytes_alg["frame"])
    roofline["other_sequence_kernel"] = {
        "kernel": "64-ch feature warp " + ("alone (separate calls)" if args.fuse_frame_warp else "+ frame warp (fused)"),
        "kernel_ms": a_ms, "algorithmic_bytes_per_launch": a_bytes,
        "achieved": a_bytes / (a_ms * 1e-3) / 1e9, "frac": a_bytes / (a_ms * 1e-3) / 1e9 / peak}
    # DRAM bytes of the same kernel from the committed ncu --set full capture (per launch)
    prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(prof) and (Hh, Ww, B) == (1088, 1920, 1) and args.flow == "smooth":
        try:
            t = json.load(open(prof))["fused" if args.fuse_frame_warp else "separate"]
            roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            roofline["traffic_source"] = t["source"]
        except Exception:
            pass

    # ---- end to end through the host-buffer API (pinned host inputs, H2D + D2H timed)
    e2e = None
    if not args.no_e2e:
        n_e = args.e2e_steps or min(args.steps, 40)
        sess = HostSession(cpu_in, models, dev, warp_algo=algo)
        for _ in range(3):
            sess.process()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with sampler:
            n_done = 0
            e0.record()
            t0 = time.perf_counter()
            while n_done < n_e or (time.perf_counter() - t0 < MIN_TIMED_S and n_done < 100 * n_e):
                sess.process()
                n_done += 1
            sess.drain()
            e1.record()
            barrier()
        e_ms = shard.max_over_ranks(e0.elapsed_time(e1), dev)
        n_done = int(shard.sum_over_ranks(n_done, dev))
        e2e = {"value": n_done / (e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": sess.h2d_bytes, "d2h_bytes_per_step": sess.d2h_bytes,
               "steps": n_done // world, "ms_per_step": e_ms / (n_done / world),
               "api": "deepsvc_b200.hotpath.HostSession.process (pinned host tensors in / out)",
               "cpu_affinity": "GPU-local NUMA node (NVML)" if numa_bound else "inherited"}
        del sess
        # the same API with the codec state the reference keeps on the device left there
        # (test_video.py:368-369: ref_frame / feature are device tensors from frame to frame)
        sess = HostSession(cpu_in, models, dev, warp_algo=algo, carry_on_device=True)
        for _ in range(3):
            sess.process()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_c = 0
        c0.record()
        t0 = time.perf_counter()
        while n_c < 4 * n_e or (time.perf_counter() - t0 < MIN_TIMED_S and n_c < 400 * n_e):
            sess.process()
            n_c += 1
        sess.drain()
        c1.record()
        barrier()
        c_ms = shard.max_over_ranks(c0.elapsed_time(c1), dev)
        n_c = int(shard.sum_over_ranks(n_c, dev))
        e2e["device_resident_state"] = {
            "value": n_c / (c_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": sess.h2d_bytes,
            "d2h_bytes_per_step": sess.d2h_bytes, "steps": n_c // world,
            "what": "HostSession(carry_on_device=True): ref_frame / feature uploaded once and warped_feature left on the "
                    "device, as the reference keeps them (test_video.py:368-369); every other tensor crosses the host per frame"}
        del sess

    # ---- baselines beside it (rank 0, N=1 only)
    cpu_baseline = cpu_1t = stock_gpu = dropin_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import reference_ops as R   # the checker, timed as the reported baselines
        models_o = oracle_models(models)
        # (a) the oracle's ops moved .cuda(): the reference's own CUDA branch (modules.py:44-62
        #     F.grid_sample + eager compressai-equivalent chain), eager, the way the reference runs
        models_og = {k: (eb.to(dev), gc.to(dev)) for k, (eb, gc) in oracle_models(models).items()}
        with torch.no_grad():
            sg_ms = time_gpu_eager(lambda: R.pframe_hotpath(gpu_in, models_og), dev, steps=10, warmup=3)
            # (b) the same calls through this package's public drop-in API, eager, no graph
            de_ms = time_gpu_eager(lambda: pframe_eager(gpu_in, models), dev, steps=20, warmup=3)
            t0 = time.perf_counter()
            for _ in range(20):
                pframe_eager(gpu_in, models)
            de_host_ms = (time.perf_counter() - t0) / 20 * 1e3    # host time to ISSUE a frame (no sync)
            torch.cuda.synchronize(dev)
        stock_gpu = {"value": 1e3 / sg_ms, "unit": UNIT, "ms_per_step": sg_ms, "steps": 10,
                     "kind": "oracle ops on cuda:0 = the reference's CUDA branch (modules.py:44-62 grid_sample "
                             "+ eager compressai-equivalent entropy chain + torch log/sum), eager, CUDA events"}
        dropin_eager = {"value": 1e3 / de_ms, "unit": UNIT, "ms_per_step": de_ms, "host_issue_ms_per_step": de_host_ms,
                        "steps": 20, "vs_stock_gpu": sg_ms / de_ms,
                        "kind": "deepsvc_b200.hotpath.pframe_eager: the drop-in ops called eagerly in "
                                "DeepSVC.forward's order (no pre-bound launches, no CUDA graph)"}
        threads = os.cpu_count() or 1
        fps, desc, _ = time_cpu_path(cpu_in, models_o, steps=3, warmup=1, threads=threads, budget_s=20.0)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc}
        # the reference's own evaluation setting: torch.set_num_threads(1) (test_video.py:16)
        fps1, desc1, _ = time_cpu_path(cpu_in, models_o, steps=2, warmup=1, threads=1, budget_s=12.0)
        cpu_1t = {"value": fps1, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc1}
        torch.set_num_threads(threads)

    if rank == 0:
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # `config` holds what defines the workload (identical in the --impl reference line); how this
            # arm ran it is under `run`
            "config": {"workload": workload, "flow": args.flow, "l2": l2_note(bytes_alg["total"], nsets)},
            "run": {"warp_algo": args.algo,
                       "per_gpu": "each rank codes its own independent sequence (no data-path collective)",
                       "timed_region": f"{len(blocks)} blocks of {args.steps} steps (>= {MIN_TIMED_S} s in total); "
                                       "value = steps / median block time, max over ranks",
                       "launch": (f"CUDA graph replay of the frame's {hp.n_launches} hot-path launches, "
                                  + {"serial": "serial order of DeepSVC.forward",
                                     "wide": "captured as their data-dependency DAG (one branch per launch: no "
                                             "op of the path consumes another's output; joined by bits_finalize)",
                                     "branches4": "captured as their data-dependency DAG (4 branches: feature warp | "
                                                  "3-ch warps | mv entropy | res entropy, joined by bits_finalize)"}[main_mode]),
                       "serial_ms_per_step": order_ms["serial"], "serial_value": world * 1e3 / order_ms["serial"],
                       "dag_ms_per_step": order_ms.get("wide", order_ms.get("branches4")),
                       "fuse_frame_warp": bool(args.fuse_frame_warp),
                       ("separate_calls_25_launches" if args.fuse_frame_warp else "fused_frame_warp_24_launches"):
                           {"dag_ms_per_step": alt_ms["dag"], "dag_value": world * 1e3 / alt_ms["dag"],
                            "serial_ms_per_step": alt_ms["serial"], "serial_value": world * 1e3 / alt_ms["serial"]},
                       "bpp_check": bpp},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_1thread": cpu_1t,
            "stock_gpu": stock_gpu, "dropin_eager": dropin_eager, "e2e": e2e,
            "gpu_launches": hp.n_launches * args.steps * len(blocks), "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier(device_ids=[local_rank])
        torch.distributed.destroy_process_group()


def run_cfg3(args):
    """Secondary line: BASELINE configs[2], the training frame-step (forward + backward) through the
    drop-in API + torch autograd, replayed as one CUDA graph.  Not the headline metric."""
    import torch
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import _lib, shard, synthetic
    from deepsvc_b200.trainstep import TrainStepHotPath, make_cotangents, train_algorithmic_bytes
    from deepsvc_b200.warp import warp_backward

    rank, local_rank, world = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for --impl ours)")
    _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    Bt, Ht, Wt = 8, 256, 256
    cpu_in = synthetic.make_pframe_inputs(B=Bt, H=Ht, W=Wt, seed=16 + rank, training=True)
    models = build_models(dev)
    for eb, gc in models.values():
        eb.train(), gc.train()
    ts = TrainStepHotPath(synthetic.to_device(cpu_in, dev), models, synthetic.to_device(make_cotangents(cpu_in), dev))
    ts.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # working set 0.8 GB > L2, flushed anyway
    # N > 1: data-parallel training.  The path's own parameters are the two bottlenecks' (the conv
    # transforms are outside it): their real gradients live in flat buckets and are all-reduced
    # after every step (mean, clamp after the reduction).  The whole-model version -- DeepSVC's
    # 82.7 MB of gradients, buckets launched from autograd hooks while backward is still running --
    # is `--workload dropin --train` (the unmodified reference model).
    grads = None
    if world > 1:
        graph_grads = [p_.grad for p_ in ts.params]   # the captured step writes these (graph-pool) tensors
        grads = shard.FlatGradBuckets(ts.params, overlap=False)   # a replayed CUDA graph fires no autograd hooks
        views = [p_.grad for p_ in ts.params]
    def step():
        ts.replay()
        if grads is not None:
            torch._foreach_copy_(views, graph_grads)   # one multi-tensor launch: the step's gradients into the buckets
            grads.allreduce(clamp=1.0)

    def barrier():
        if world > 1:
            torch.distributed.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = shard.max_over_ranks(ev0.elapsed_time(ev1), dev)
    nb = train_algorithmic_bytes(Bt, Ht, Wt)
    # dominant kernel: the 64-ch warp backward (both gradients), CUDA events, L2 flushed
    x, f = ts.inp["feature"].detach(), ts.inp["flow"].detach()
    g = torch.randn_like(x)
    gin = torch.empty_like(x)
    ts_k = []
    lin = dsvc.warp._base_grids(dev, Ht, Wt)
    sc = dsvc.warp._scales(Ht, Wt)
    gflow = torch.empty_like(f)
    ws = dsvc.warp._bwd_workspace(dev, Bt, Ht, Wt, big=True)
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.load().dsvc_warp_bwd_ws_f32(g.data_ptr(), x.data_ptr(), f.data_ptr(), gin.data_ptr(),
                                                    gflow.data_ptr(), Bt, 64, Ht, Wt, lin[0].data_ptr(), lin[1].data_ptr(),
                                                    sc[0], sc[1], sc[2], sc[3], _lib.FLOW_MUL_RECIPROCAL, _lib.LAYOUT_NCHW,
                                                    ws.data_ptr(), ws.numel(),
                                                    torch.cuda.current_stream(dev).cuda_stream), "dsvc_warp_bwd_ws_f32")
        e1.record()
        ts_k.append((e0, e1))
    torch.cuda.synchronize(dev)
    k_ms = statistics.median(a.elapsed_time(b) for a, b in ts_k)
    peak, peak_src = measured_peak()
    ach = nb["feature_bwd"] / (k_ms * 1e-3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "cfg3 training frame-steps/sec (warp+entropy path, forward+backward)",
            "value": world * args.steps / (ms * 1e-3), "unit": "frame-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3: B=8 256x256 crops, noise-mode entropy models, forward + backward of 6 warps "
                                   "(64-ch: both gradients; 3-ch: flow only) + 16 GC + 2 EB + log-likelihood sums through "
                                   "deepsvc_b200's drop-in ops and torch autograd, one CUDA graph per step",
                       "algorithmic_bytes_per_step": nb["total"], "l2": "0.8 GB working set per step vs 126 MB L2",
                       "allreduce": ("none (N = 1)" if world == 1 else
                                     "the path's own parameters (two EntropyBottlenecks): real p.grad views in flat buckets, NCCL "
                                     "all-reduce + mean + clamp inside the timed region; the whole-model all-reduce overlapped with "
                                     "backward is `--workload dropin --train`")},
            "roofline": {"bound": "hbm", "kernel": "warp_bwd_cell_staged (64-ch, both gradients, 8x256x256)", "achieved": ach,
                         "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": None,
                         "algorithmic_bytes_per_launch": nb["feature_bwd"], "kernel_ms": k_ms,
                         "note": "whole dsvc_warp_bwd_ws_f32 call (table memsets + cell_build + cell-order kernel + overflow launch), L2 flushed",
                         "whole_step": {"algorithmic_bytes": nb["total"],
                                        "achieved_gbs": nb["total"] * args.steps / (ms * 1e-3) / 1e9}},
            "cpu_baseline": None, "e2e": None, "clocks": sampler.summary()}), flush=True)
    if world > 1:
        torch.distributed.barrier(device_ids=[local_rank])
        torch.distributed.destroy_process_group()


def run_cfg4(args):
    """Secondary line: BASELINE configs[3] -- 7 independent 1080p sequences of 96 frames, cut into
    (sequence, GOP) jobs and assigned to the ranks by shard.assign_jobs (static LPT, no data-path
    collective).  STRONG scaling: the job list is fixed, every rank replays the frame graph once
    per P-frame of its jobs; time = max over ranks on the device."""
    import torch
    from deepsvc_b200 import _lib, shard, synthetic
    from deepsvc_b200.hotpath import PFrameHotPath
    rank, local_rank, world = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for --impl ours)")
    _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    jobs = shard.make_gop_jobs([96] * 7, args.gop)
    assignment = shard.assign_jobs(jobs, world)
    mine = assignment[rank]
    cpu_in = synthetic.make_pframe_inputs(B=B, H=H, W=W, seed=16 + rank)
    hp = PFrameHotPath(synthetic.to_device(cpu_in, dev), build_models(dev), fuse_frame_warp=args.fuse_frame_warp)
    hp.capture()
    for _ in range(10):
        hp.replay()
    if world > 1:
        torch.distributed.barrier(device_ids=[local_rank])
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with sampler:
        ev0.record()
        for job in mine:                 # GOPs are independent; P-frames inside one are serial
            for _ in range(job.p_frames):
                hp.replay()
        ev1.record()
        torch.cuda.synchronize(dev)
    ms_local = ev0.elapsed_time(ev1)
    ms = shard.max_over_ranks(ms_local, dev)
    total = sum(j.p_frames for j in jobs)
    if rank == 0:
        print(json.dumps({
            "metric": "1080p P-frames/sec (warp+entropy path), 7 x 96-frame sequences sharded by GOP",
            "value": total / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": total, "warmup": 10,
            "ms_per_step": ms / total, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"cfg4: 7 sequences x 96 frames at 1920x1088, GOP {args.gop}: {len(jobs)} (sequence, GOP) "
                                   f"jobs = {total} P-frames, static LPT assignment, no data-path collective",
                       "jobs_per_rank": [len(a) for a in assignment],
                       "p_frames_per_rank": [sum(j.p_frames for j in a) for a in assignment],
                       "balance": shard.balance(assignment)},
            "clocks": sampler.summary()}), flush=True)
    if world > 1:
        torch.distributed.barrier(device_ids=[local_rank])
        torch.distributed.destroy_process_group()


def run_codec(args):
    """Secondary line: the symbol pipeline of coded 1080p P-frames --
    image_model.py:201-257 for both codecs: per frame 18 fused launches write every slice's symbols
    and table indexes into one device buffer, ONE pinned device-to-host copy per frame, and the
    frame's four rANS streams (mv y, mv z, res y, res z; compressai's wire format) are coded on a
    pool of host threads together with those of the other frames in flight."""
    import torch
    import deepsvc_b200 as dsvc  # noqa: F401
    from deepsvc_b200 import _lib, ans, synthetic
    from deepsvc_b200.codec import FrameSymbolDecoder, FrameSymbolPipeline
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for --impl ours)")
    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    Hh, Ww = frame_size(args)
    cpu_in = synthetic.make_pframe_inputs(B=1, H=Hh, W=Ww, seed=16)
    d = synthetic.to_device(cpu_in, dev)
    models = build_models(dev)
    for eb, gc in models.values():
        eb.update(force=True)
        gc.update_scale_table(synthetic.get_scale_table().to(dev))
    nsym = sum(d[f"{n}_y"].numel() + d[f"{n}_z"].numel() for n in ("mv", "res"))
    threads = args.threads or (os.cpu_count() or 1)
    group = max(2, threads // 4)                     # frames coded together: 4 streams each
    pipe = FrameSymbolPipeline(d, models, depth=2 * group)

    def encode_frames(n_frames):
        """Double-buffered groups: while the host codes group k, the device produces group k + 1."""
        streams, nbytes = None, 0
        pending = None
        done = 0
        g = 0
        while done < n_frames or pending is not None:
            cur = None
            if done < n_frames:
                k = min(group, n_frames - done)
                base = (g % 2) * group
                for j in range(k):
                    pipe.launch(base + j, d)
                cur = (base, k)
                done += k
                g += 1
            if pending is not None:
                jobs = [job for j in range(pending[1]) for job in pipe.jobs(pending[0] + j)]
                out = ans.encode_many(jobs, threads)
                streams = out[:4]
                nbytes = sum(len(b) for b in out[:4])
            pending = cur
        return streams, nbytes

    encode_frames(2 * group)                          # warm-up
    steps = max(args.steps if args.steps != 1000 else 400, 4 * group)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    streams, nbytes = encode_frames(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0

    # ---- decode: the same streams back to y_hat / z_hat on the device, `group` frames per call
    dec = [FrameSymbolDecoder(pipe, d) for _ in range(group)]
    idx_jobs = [(j[1].copy(), j[2]) for j in pipe.jobs(0)]
    y_ref, z_ref = pipe.reconstruction(0)
    got = dec[0].decode(streams, idx_jobs, threads)
    torch.cuda.synchronize()
    roundtrip = all(torch.equal(got[n][0], y_ref[n]) and torch.equal(got[n][1], z_ref[n]) for n in ("mv", "res"))

    def decode_frames(n_frames):
        from concurrent.futures import ThreadPoolExecutor
        # (decode_many releases the GIL inside ctypes: the frames of a group decode concurrently)
        with ThreadPoolExecutor(max_workers=group) as pool:
            for f0 in range(0, n_frames, group):
                k = min(group, n_frames - f0)
                list(pool.map(lambda j: dec[j].decode(streams, idx_jobs, max(1, threads // group)), range(k)))

    decode_frames(group)
    dsteps = max(steps // 2, 2 * group)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    decode_frames(dsteps)
    torch.cuda.synchronize()
    dt_dec = time.perf_counter() - t0
    print(json.dumps({
        "metric": "coded 1080p P-frames/sec (symbol pipeline: quantise+index on GPU, one pinned D2H per frame, "
                  "rANS streams on host threads)",
        "value": steps / dt, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": 2 * group,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{Ww}x{Hh} P-frame, both codecs, {nsym} symbols per frame in 4 rANS streams "
                               f"(mv y / mv z / res y / res z); {pipe.n_launches} launches + 1 pinned D2H per frame; "
                               f"{group} frames per coder call on {threads} host threads (of {os.cpu_count()}); "
                               "double-buffered against the device",
                   "bytes_per_frame": nbytes, "d2h_bytes_per_frame": 8 * pipe.n_total,
                   "timing": "host wall clock around synchronised runs (the host coder is the bound)",
                   "msymbols_per_s": nsym * steps / dt / 1e6,
                   "decode": {"value": dsteps / dt_dec, "unit": "frames/s", "ms_per_step": dt_dec / dsteps * 1e3,
                              "msymbols_per_s": nsym * dsteps / dt_dec / 1e6, "round_trip_exact": roundtrip,
                              "what": "4 streams per frame -> host threads -> one pinned upload -> dequantise on the device"}}}),
          flush=True)


def run_dropin(args):
    """Secondary line: the UNMODIFIED reference ``DeepSVC`` (staged copy under the git-ignored
    ``oracle/_ref``; conv transforms included) timed on the GPU stock -- torch ``grid_sample`` +
    the oracle's eager compressai shim, i.e. the reference's own CUDA path -- and through
    ``patch_reference()`` + ``swap_entropy_models()``.  ``--train``: one training step
    (``Learner.py:1306-1343`` shape: B=8 256x256 crops, forward + backward + gradient sync) with the
    data-parallel all-reduce of the model's real ``p.grad`` views overlapped with backward
    (``shard.FlatGradBuckets``)."""
    import torch
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import _lib, shard
    from oracle import stage_reference   # baseline leg: the reference itself
    rank, local_rank, world = shard.init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    if not stage_reference.available():
        if rank == 0:
            print(json.dumps({"workload": "dropin", "unavailable": "oracle/_ref not staged (run build() where /root/reference is mounted)"}))
        return
    _lib.load()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    modules, image_model, video_model = stage_reference.import_reference()
    torch.manual_seed(16)
    model = video_model.DeepSVC().to(dev)
    train = args.train
    Bt, Hh, Ww = (8, 256, 256) if train else (1,) + tuple(frame_size(args))
    g = torch.Generator().manual_seed(3 + rank)
    ref = torch.rand(Bt, 3, Hh, Ww, generator=g).to(dev)
    cur = (torch.roll(ref.cpu(), shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(Bt, 3, Hh, Ww, generator=g)).clamp(0, 1).to(dev)
    sm = torch.rand(Bt, 256, Hh // 4, Ww // 4, generator=g).to(dev)
    fea = (torch.randn(Bt, 64, Hh, Ww, generator=g) * 0.5).to(dev)
    steps = min(args.steps, 30 if not train else 20)
    buckets = None

    def fwd():
        with torch.no_grad():
            return model(ref, cur, sm, fea)

    def train_step():
        for p_ in model.parameters():
            p_.grad = None                      # optimizer.zero_grad(set_to_none=True), Learner.py:177
        out = model(ref, cur, sm, fea)
        loss = out[2] * 2048 + out[7] + model.aux_loss() * 0.0   # lambda * mse + bpp (Learner.py:1343 shape)
        loss.backward()
        if buckets is not None:
            buckets.allreduce(clamp=1.0)        # waits the overlapped buckets; mean; clamp (Learner.py:1687-1691)

    step = train_step if train else fwd
    model.train(train)

    def timed():
        ms = time_gpu_eager(step, dev, steps=steps, warmup=3)
        return shard.max_over_ranks(ms, dev)

    def host_issue():
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        t = (time.perf_counter() - t0) / steps * 1e3
        torch.cuda.synchronize(dev)
        return t

    if world > 1 and train:
        buckets = shard.FlatGradBuckets(list(model.parameters()))
    stock_ms = timed()
    stock_host = host_issue()
    if buckets is not None:
        for h in buckets._hooks:
            h.remove()
    dsvc.patch_reference(modules, video_model, image_model)
    n_swapped = dsvc.swap_entropy_models(model)
    if world > 1 and train:
        buckets = shard.FlatGradBuckets(list(model.parameters()))
    patched_ms = timed()
    patched_host = host_issue()
    dsvc.unpatch_reference()
    if rank == 0:
        unit = "frame-steps/s" if train else "frames/s"
        print(json.dumps({
            "metric": ("reference DeepSVC training steps/sec (whole model, B=8 256x256)" if train else
                       f"reference DeepSVC.forward frames/sec (whole model incl. conv transforms, {Ww}x{Hh})"),
            "value": world * 1e3 / patched_ms, "unit": unit, "n_gpus": world, "steps": steps, "warmup": 3,
            "ms_per_step": patched_ms, "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("unmodified reference video_model.DeepSVC (oracle/_ref), eager, random-init weights; "
                                    + ("training step: forward + backward" + (" + FlatGradBuckets all-reduce of the model's "
                                       f"{sum(p.numel() for p in model.parameters()) * 4 / 1e6:.1f} MB of gradients overlapped "
                                       "with backward, mean + clamp" if world > 1 else "") if train else "inference forward")),
                       "entropy_modules_swapped": n_swapped},
            "stock_gpu": {"value": world * 1e3 / stock_ms, "unit": unit, "ms_per_step": stock_ms, "host_issue_ms": stock_host,
                          "kind": "the same model unpatched: torch grid_sample + eager compressai shim"},
            "patched": {"value": world * 1e3 / patched_ms, "unit": unit, "ms_per_step": patched_ms, "host_issue_ms": patched_host},
            "speedup_vs_stock_gpu": stock_ms / patched_ms}), flush=True)
    if world > 1:
        torch.distributed.barrier(device_ids=[local_rank])
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port",
               os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg3":
        run_cfg3(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    elif args.workload == "codec":
        run_codec(args)
    elif args.workload == "dropin":
        run_dropin(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
