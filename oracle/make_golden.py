"""ORACLE / TEST INFRASTRUCTURE ONLY.  Generates ``tests/golden/*.npz``.

Run in the build container (where ``/root/reference`` is mounted):

    python oracle/make_golden.py

It imports the UNMODIFIED reference (``/root/reference/{modules,image_model,
video_model}.py``) with ``oracle/shim`` standing in for the absent third-party
packages, and records:

* ``warp_*.npz``      inputs and outputs of the reference's own ``modules.torch_warp``
                      (CPU branch, ``modules.py:26-43``) -- this pins the warp oracle;
* ``callsite_*.npz``  the tensors that reach every GaussianConditional /
                      EntropyBottleneck call site inside ``DeepSVC.forward``
                      (``image_model.py:155,181``; captured with forward hooks on a
                      seeded random-init model) together with what the calls returned,
                      plus the model's warps and bpp -- this anchors the entropy oracle
                      on the reference's call sites (the arithmetic itself is the shim's:
                      PARITY UNPINNED, see oracle/shim/compressai/__init__.py);
* ``entropy_kat.npz`` hand-made known-answer cases (round-half-even ties, scales on the
                      table entries and on the 0.11 bound, likelihood floor).

The GPU box has no ``/root/reference``; tests there read only these fixtures.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "tests", "golden")


def warp_cases():
    import modules  # the reference
    from deepsvc_b200 import synthetic

    gen = torch.Generator().manual_seed(16)
    cases = {
        "smooth_b2c3_24x40": (2, 3, 24, 40, "smooth"),
        "stress_b1c5_17x23": (1, 5, 17, 23, "stress"),   # odd sizes: linspace midpoint branch
        "border_b1c4_64x64": (1, 4, 64, 64, "border"),
        "smooth_b1c64_32x48": (1, 64, 32, 48, "smooth"),
    }
    for name, (B, C, H, W, kind) in cases.items():
        inp = torch.randn(B, C, H, W, generator=gen)
        if kind == "smooth":
            flow = synthetic.smooth_flow(B, H, W, gen)
        elif kind == "stress":
            flow = synthetic.stress_flow(B, H, W, gen, sigma=6.0)
        else:
            flow = synthetic.border_flow(B, H, W, gen, margin=16, reach=24.0)
        inp.requires_grad_(True)
        flow.requires_grad_(True)
        out = modules.torch_warp(inp, flow)
        gout = torch.randn(out.shape, generator=gen)
        gin, gflow = torch.autograd.grad(out, (inp, flow), gout)
        np.savez_compressed(os.path.join(OUT, f"warp_{name}.npz"), input=inp.detach().numpy(),
                            flow=flow.detach().numpy(), out=out.detach().numpy(),
                            grad_out=gout.numpy(), grad_input=gin.numpy(), grad_flow=gflow.numpy())
        print("warp", name, tuple(out.shape))


def callsite_case():
    import video_model  # the reference

    torch.manual_seed(16)
    model = video_model.DeepSVC().eval()
    # exercise the tanh gate and non-zero medians of the factorised prior
    with torch.no_grad():
        for codec in (model.mv_codec, model.res_codec):
            eb = codec.entropy_bottleneck
            for i in range(4):
                getattr(eb, f"_factor{i}").normal_(0, 0.1)
            eb.quantiles[:, 0, 1].normal_(0, 0.3)
    rec = {}

    def hook_gc(tag):
        calls = []

        def fn(mod, args, kwargs, output):
            x, scales, means = args[0], args[1], args[2]
            calls.append((x.detach().clone(), scales.detach().clone(), means.detach().clone(),
                          output[0].detach().clone(), output[1].detach().clone()))
        rec[tag] = calls
        return fn

    def hook_eb(tag):
        calls = []

        def fn(mod, args, kwargs, output):
            calls.append((args[0].detach().clone(), output[0].detach().clone(),
                          output[1].detach().clone()))
        rec[tag] = calls
        return fn

    for cname, codec in (("mv", model.mv_codec), ("res", model.res_codec)):
        codec.gaussian_conditional.register_forward_hook(hook_gc(f"gc_{cname}"), with_kwargs=True)
        codec.entropy_bottleneck.register_forward_hook(hook_eb(f"eb_{cname}"), with_kwargs=True)

    warps = []
    import modules
    orig = modules.torch_warp

    def rec_warp(a, b):
        o = orig(a, b)
        warps.append((a.detach().clone(), b.detach().clone(), o.detach().clone()))
        return o
    modules.torch_warp = rec_warp
    video_model.torch_warp = rec_warp

    ref = torch.rand(1, 3, 64, 128)
    cur = (ref + 0.05 * torch.randn(1, 3, 64, 128)).clamp(0, 1)
    sm = torch.rand(1, 256, 16, 32)
    with torch.no_grad():
        out = model(ref, cur, sm, None)
    modules.torch_warp = orig
    video_model.torch_warp = orig

    d = {}
    for cname, codec in (("mv", model.mv_codec), ("res", model.res_codec)):
        eb = codec.entropy_bottleneck
        for n, p in eb.named_parameters():
            d[f"eb_{cname}_param_{n}"] = p.detach().numpy()
        (z, z_out, z_lik), = rec[f"eb_{cname}"]
        d[f"eb_{cname}_z"], d[f"eb_{cname}_out"], d[f"eb_{cname}_lik"] = z.numpy(), z_out.numpy(), z_lik.numpy()
        gcs = rec[f"gc_{cname}"]
        assert len(gcs) == 8
        for k, key in enumerate(("x", "scales", "means", "out", "lik")):
            d[f"gc_{cname}_{key}"] = np.stack([c[k].numpy() for c in gcs])
    assert len(warps) == 6
    for i, (a, b, o) in enumerate(warps):
        d[f"warp{i}_input"], d[f"warp{i}_flow"], d[f"warp{i}_out"] = a.numpy(), b.numpy(), o.numpy()
    d["bpp_res"], d["bpp_mv"], d["bpp"] = (float(out[5]), float(out[6]), float(out[7]))
    d["scale_table"] = model.mv_codec.gaussian_conditional.scale_table.numpy() \
        if model.mv_codec.gaussian_conditional.scale_table.numel() else np.zeros(0, np.float32)
    np.savez_compressed(os.path.join(OUT, "callsite_deepsvc_64x128.npz"), **d)
    print("callsite: bpp", d["bpp"], "gc calls", len(rec["gc_mv"]) + len(rec["gc_res"]))


def entropy_kat():
    from oracle import reference_ops as R

    eb, gc = R.make_entropy_models(8, seed=3)
    table = gc.scale_table.clone()
    mu = torch.tensor([0.0, 0.25, -1.5, 3.0, 0.1, -0.3, 7.0, 0.0])
    # ties: y - mu exactly k + 0.5 -> half-to-even
    ties = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 3.5, -3.5])
    y_t = mu + ties
    # scales exactly on table entries, on / below the bound, and between entries
    sc = torch.cat([table[:4], torch.tensor([0.11, 0.05, 0.1100001, 300.0])])
    # far tails -> likelihood floor
    y_f = mu + torch.tensor([40.0, -40.0, 12.0, -12.0, 5.0, -5.0, 0.0, 0.0]) * sc.clamp(min=0.11)
    x = torch.stack([y_t, y_f, mu + torch.linspace(-2, 2, 8)]).reshape(1, 3, 1, 8)
    scales = torch.stack([sc, sc, table[20:28]]).reshape(1, 3, 1, 8)
    means = mu.repeat(3).reshape(1, 3, 1, 8)
    gc.eval()
    out, lik = gc(x, scales, means)
    sym = gc.quantize(x, "symbols", means)
    idx = gc.build_indexes(scales)
    full_idx = gc.build_indexes(table.reshape(1, 1, 1, -1))
    np.savez_compressed(os.path.join(OUT, "entropy_kat.npz"), x=x.numpy(), scales=scales.numpy(),
                        means=means.numpy(), out=out.numpy(), lik=lik.numpy(), symbols=sym.numpy(),
                        indexes=idx.numpy(), table=table.numpy(), table_indexes=full_idx.numpy())
    print("kat: symbols", sym.flatten()[:8].tolist(), "idx", idx.flatten()[:8].tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    warp_cases()
    callsite_case()
    entropy_kat()
