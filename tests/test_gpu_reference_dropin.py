"""The UNMODIFIED reference model on the B200, stock vs through the drop-ins.

``oracle/stage_reference.py`` stages byte-identical copies of the reference's
``modules.py`` / ``image_model.py`` / ``video_model.py`` into the git-ignored ``oracle/_ref``
(the GPU box has no ``/root/reference``).  With the oracle's compressai shim on ``sys.path``
they import unmodified, and ``video_model.DeepSVC`` runs

  * stock:   torch CUDA ``F.grid_sample`` (``modules.py:44-62``) + eager shim entropy models;
  * patched: ``deepsvc_b200.patch_reference()`` + ``swap_entropy_models(model)`` (same weights).

Three levels of parity (north_star tolerances, written out below):

1. call sites -- every hot-path call the stock ``DeepSVC.forward`` (``video_model.py:27-71``)
   makes is recorded (inputs and outputs) and replayed through the drop-in op on the SAME
   inputs: warps <= 1e-5 relative, quantised latents bit-exact, likelihood bit sums <= 1e-4;
2. end to end -- the patched model's outputs against the stock model's (bpp <= 1e-4 relative);
3. ``compress`` -> ``decompress`` (``video_model.py:137-167``) round trip through the drop-in
   symbol pipeline + C++ range coder, and the bit streams against the stock run's.

The measured deviations are also written to ``gpurun_out/dropin_parity_*.json``.
"""
import json
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = [(256, 448), (1088, 1920)]


def _dev():
    return torch.device("cuda:0")


@pytest.fixture(autouse=True)
def _deterministic_convs():
    """The conv transforms are outside the path; make them reproducible so that any
    stock-vs-patched difference comes from the drop-in ops."""
    old = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark,
           torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark,
     torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32) = old


def _make_model(video_model, seed=16):
    """Seeded random-init DeepSVC (``utils.py:16`` default seed) with the bottleneck gates and
    medians perturbed (a fresh model has ``_factor == 0``: the tanh terms would vanish)."""
    torch.manual_seed(seed)
    m = video_model.DeepSVC()
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for codec in (m.mv_codec, m.res_codec):
            eb = codec.entropy_bottleneck
            for i in range(4):
                f = getattr(eb, f"_factor{i}")
                f.copy_(torch.randn(f.shape, generator=g) * 0.1)
            eb.quantiles[:, 0, 1] = torch.randn(eb.quantiles.size(0), generator=g) * 0.3
    return m.to(_dev()).eval()


def _make_inputs(H, W, seed=3, with_feature=True):
    g = torch.Generator().manual_seed(seed)
    ref = torch.rand(1, 3, H, W, generator=g)
    # the current frame is the reference moved by a few pixels plus noise: a non-trivial flow
    cur = (torch.roll(ref, shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(1, 3, H, W, generator=g)).clamp(0, 1)
    sm = torch.rand(1, 256, H // 4, W // 4, generator=g)
    fea = torch.randn(1, 64, H, W, generator=g) * 0.5 if with_feature else None
    dev = _dev()
    return ref.to(dev), cur.to(dev), sm.to(dev), (fea.to(dev) if fea is not None else None)


class _Recorder:
    """Records what reaches / leaves every hot-path call site of one forward."""

    def __init__(self, modules, video_model, model):
        self.mods = (modules, video_model)
        self.model = model
        self.warps, self.gc, self.eb = [], [], []
        self._hooks, self._orig = [], {}

    def __enter__(self):
        for mod in self.mods:
            orig = mod.torch_warp
            self._orig[mod] = orig

            def rec(inp, flow, _orig=orig):
                out = _orig(inp, flow)
                self.warps.append((inp.detach().clone(), flow.detach().clone(), out.detach().clone()))
                return out
            mod.torch_warp = rec
        for name, codec in (("mv", self.model.mv_codec), ("res", self.model.res_codec)):
            def gc_hook(_m, args, out, name=name):
                self.gc.append((name, tuple(a.detach().clone() for a in args), tuple(o.detach().clone() for o in out)))

            def eb_hook(_m, args, out, name=name):
                self.eb.append((name, args[0].detach().clone(), tuple(o.detach().clone() for o in out)))
            self._hooks.append(codec.gaussian_conditional.register_forward_hook(gc_hook))
            self._hooks.append(codec.entropy_bottleneck.register_forward_hook(eb_hook))
        return self

    def __exit__(self, *a):
        for mod, orig in self._orig.items():
            mod.torch_warp = orig
        for h in self._hooks:
            h.remove()


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-30)


def _dump(name, obj):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", name), "w") as fh:
        json.dump(obj, fh, indent=1)


@pytest.mark.parametrize("size", SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_callsites_of_stock_forward_replayed_through_dropins(reference_modules, size):
    import deepsvc_b200 as d
    from deepsvc_b200.patch import _convert_eb, _convert_gc
    modules, image_model, video_model = reference_modules
    H, W = size
    model = _make_model(video_model)
    ref, cur, sm, fea = _make_inputs(H, W)
    with torch.no_grad(), _Recorder(modules, video_model, model) as rec:
        model(ref, cur, sm, fea)
    # DeepSVC.forward: 4 SpyNet warps + frame warp + feature warp, 16 GC slice calls, 2 EB calls
    assert len(rec.warps) == 6 and len(rec.gc) == 16 and len(rec.eb) == 2
    report = {"size": [H, W], "warps": [], "gc": [], "eb": []}
    for inp, flow, out in rec.warps:
        got = d.torch_warp(inp, flow)
        err = (got - out).abs().max().item()
        tol = 1e-5 * max(1.0, out.abs().max().item())
        report["warps"].append({"shape": list(inp.shape), "max_abs_err": err,
                                "bit_identical_frac": (got == out).float().mean().item()})
        assert err <= tol, (tuple(inp.shape), err, tol)
        assert torch.allclose(got, out, rtol=1e-5, atol=2e-6)
    for name, codec in (("mv", model.mv_codec), ("res", model.res_codec)):
        gc = _convert_gc(codec.gaussian_conditional)
        eb = _convert_eb(codec.entropy_bottleneck)
        with torch.no_grad():
            for n, (y, s, mu), (out, lik) in [r for r in rec.gc if r[0] == name]:
                g_out, g_lik = gc(y, s, mu)
                y_hat, _, parts = gc.forward_fused(y, s, mu)
                assert torch.equal(g_out, out), f"{name}: quantised outputs not bit-exact"
                assert torch.equal(y_hat, image_model.ste_round(y - mu) + mu), f"{name}: y_hat (image_model.py:183)"
                assert torch.allclose(g_lik, lik, rtol=2e-4, atol=1e-12)
                bits_ref = torch.log(lik.double()).sum().item()
                r = _rel(parts.sum().item(), bits_ref)
                report["gc"].append({"codec": name, "shape": list(y.shape), "ln_lik_rel_err": r,
                                     "lik_max_rel": ((g_lik - lik).abs() / lik).max().item()})
                assert r <= 1e-4
            for n, z, (out, lik) in [r for r in rec.eb if r[0] == name]:
                e_out, e_lik = eb(z)
                z_hat, _, parts = eb.forward_fused(z)
                assert torch.equal(e_out, out), f"{name}: z outputs not bit-exact"
                med = codec.entropy_bottleneck._get_medians()
                assert torch.equal(z_hat, image_model.ste_round(z - med) + med)   # image_model.py:160-162
                assert torch.allclose(e_lik, lik, rtol=2e-4, atol=1e-12)
                r = _rel(parts.sum().item(), torch.log(lik.double()).sum().item())
                report["eb"].append({"codec": name, "shape": list(z.shape), "ln_lik_rel_err": r})
                assert r <= 1e-4
    _dump(f"dropin_parity_callsites_{H}x{W}.json", report)


@pytest.mark.parametrize("size", SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_forward_patched_vs_stock(reference_modules, size):
    """``DeepSVC.forward`` end to end: stock, stock again (run-to-run noise floor), patched."""
    import deepsvc_b200 as d
    modules, image_model, video_model = reference_modules
    H, W = size
    model = _make_model(video_model)
    ref, cur, sm, fea = _make_inputs(H, W)
    names = ["recon_image", "feature", "mse_loss", "warp_loss", "mc_loss", "bpp_res", "bpp_mv", "bpp"]

    def run():
        with torch.no_grad(), _Recorder(modules, video_model, model) as rec:
            out = model(ref, cur, sm, fea)
        return dict(zip(names, out)), rec

    stock, rec_s = run()
    stock2, _ = run()
    try:
        d.patch_reference(modules, video_model, image_model)
        assert d.swap_entropy_models(model) == 4
        patched, rec_p = run()
    finally:
        d.unpatch_reference()
    assert modules.torch_warp is not d.torch_warp
    rep = {"size": [H, W]}
    for k in ("bpp", "bpp_mv", "bpp_res", "mse_loss", "warp_loss", "mc_loss"):
        rep[k] = {"stock": float(stock[k]), "patched": float(patched[k]), "rel": _rel(patched[k], stock[k]),
                  "stock_rerun_rel": _rel(stock2[k], stock[k])}
    # quantised latents seen by the two runs (outputs of the 16 GC calls = round(y - mu) + mu)
    tot = mism = 0
    for (_, _, (o_s, _)), (_, _, (o_p, _)) in zip(rec_s.gc, rec_p.gc):
        tot += o_s.numel()
        mism += int((o_s != o_p).sum().item())
    rep["quantised_latents"] = {"elements": tot, "mismatched": mism}
    rep["recon_max_abs_diff"] = (patched["recon_image"] - stock["recon_image"]).abs().max().item()
    rep["warped_frame_max_abs_diff"] = (rec_p.warps[4][2] - rec_s.warps[4][2]).abs().max().item()
    rep["warped_feature_max_rel_diff"] = ((rec_p.warps[5][2] - rec_s.warps[5][2]).abs().max().item()
                                          / max(1.0, rec_s.warps[5][2].abs().max().item()))
    _dump(f"dropin_parity_forward_{H}x{W}.json", rep)
    # north_star: bpp within 1e-4 relative; warped tensors 1e-5 relative.  The warps inside the
    # model see inputs that already went through conv transforms fed by earlier drop-in outputs
    # (SpyNet levels), so the end-to-end warp tolerance is 10x the per-call one.
    for k in ("bpp", "bpp_mv", "bpp_res"):
        assert rep[k]["rel"] <= 1e-4, (k, rep[k])
    assert rep["warped_frame_max_abs_diff"] <= 1e-4 and rep["warped_feature_max_rel_diff"] <= 1e-4
    assert mism <= 1e-4 * tot, rep["quantised_latents"]   # a tie may flip where a conv input moved by 1 ulp
    assert rep["recon_max_abs_diff"] <= 1e-3


def test_forward1_and_forward_msssim_patched_vs_stock(reference_modules, monkeypatch):
    """The other callers of the path in ``video_model.py``: ``forward1`` (``:73-94``, warp at ``:83``; the
    stage-1 training forward) and ``forward_msssim`` (``:96-135``, warp at ``:106``), stock vs patched;
    then ``forward1`` in training mode with a backward pass (noise-mode entropy models, the 64-ch
    feature warp's gradient with respect to the feature through the cell-order backward kernel).
    MS-SSIM itself is outside the path (the shim has no implementation): a fixed stand-in is used
    for both runs."""
    import deepsvc_b200 as d
    modules, image_model, video_model = reference_modules
    H, W = 256, 448
    model = _make_model(video_model)
    ref, cur, sm, fea = _make_inputs(H, W)
    monkeypatch.setattr(video_model, "ms_ssim", lambda a, b, data_range=1.0: 1.0 - torch.mean((a - b) ** 2))

    def run_eval():
        with torch.no_grad():
            return model.forward1(ref, cur, sm, fea), model.forward_msssim(ref, cur, sm, fea)

    def run_train(full=False):
        model.train()
        try:
            torch.manual_seed(11)
            f = fea.clone().requires_grad_(True)
            model.zero_grad(set_to_none=True)
            if full:   # DeepSVC.forward, the stage >= 4 training forward (Learner.py:1332-1343): both codecs
                recon, feature, mse_loss, warp_loss, mc_loss, bpp_res, bpp_mv, bpp = model(ref, cur, sm, f)
                (mse_loss + 0.1 * warp_loss + 0.1 * mc_loss + 0.01 * bpp).backward()
                a, b_ = mse_loss, bpp
            else:
                predict_frame, warp_loss, mc_loss, bpp_mv = model.forward1(ref, cur, sm, f)
                (mc_loss + 0.1 * warp_loss + 0.01 * bpp_mv).backward()
                a, b_ = mc_loss, bpp_mv
            pg = {n: p.grad.detach().clone() for n, p in model.named_parameters()
                  if p.grad is not None and (n.startswith("mv_codec.g_a.0") or n.startswith("opticFlow.moduleBasic.0")
                                             or n.startswith("res_codec.g_a.0"))}
            return f.grad.detach().clone(), pg, float(a.detach()), float(b_.detach())
        finally:
            model.eval()
            model.zero_grad(set_to_none=True)

    stock_e, stock_t, stock_t2 = run_eval(), run_train(), run_train()
    stock_f, stock_f2 = run_train(True), run_train(True)
    try:
        d.patch_reference(modules, video_model, image_model)
        assert d.swap_entropy_models(model) == 4
        patched_e, patched_t, patched_f = run_eval(), run_train(), run_train(True)
    finally:
        d.unpatch_reference()
    rep = {}
    # eval outputs: scalars within 1e-4 relative, images within 1e-3 absolute (conv inputs moved by ulps)
    for which, names in ((0, ["predict_frame", "warp_loss", "mc_loss", "bpp_mv"]),
                         (1, ["recon_image", "feature", "msssim", "warp_msssim", "mc_msssim", "bpp_res", "bpp_mv", "bpp"])):
        for nm, a, b in zip(names, patched_e[which], stock_e[which]):
            if a.dim() == 0:
                rep[f"{which}:{nm}"] = _rel(a, b)
                assert _rel(a, b) <= 1e-4, (nm, float(a), float(b))
            else:
                err = (a - b).abs().max().item()
                rep[f"{which}:{nm}"] = err
                assert err <= 1e-3 * max(1.0, b.abs().max().item()), (nm, err)
    # training: losses and gradients against the stock run, with the stock run-to-run noise as the floor
    def gdiff(x, y):
        return (x - y).abs().max().item() / max(y.abs().max().item(), 1e-30)
    floor = max([gdiff(stock_t2[0], stock_t[0])] + [gdiff(stock_t2[1][k], stock_t[1][k]) for k in stock_t[1]])
    tol = max(1e-3, 20 * floor)
    rep["train"] = {"mc_loss_rel": _rel(patched_t[2], stock_t[2]), "bpp_mv_rel": _rel(patched_t[3], stock_t[3]),
                    "grad_feature": gdiff(patched_t[0], stock_t[0]), "stock_rerun_floor": floor,
                    "grad_params": {k: gdiff(patched_t[1][k], stock_t[1][k]) for k in stock_t[1]}}
    floor_f = max([gdiff(stock_f2[0], stock_f[0])] + [gdiff(stock_f2[1][k], stock_f[1][k]) for k in stock_f[1]])
    tol_f = max(1e-3, 20 * floor_f)
    rep["train_full_forward"] = {"mse_loss_rel": _rel(patched_f[2], stock_f[2]), "bpp_rel": _rel(patched_f[3], stock_f[3]),
                                 "grad_feature": gdiff(patched_f[0], stock_f[0]), "stock_rerun_floor": floor_f,
                                 "grad_params_max": max(gdiff(patched_f[1][k], stock_f[1][k]) for k in stock_f[1])}
    _dump("dropin_parity_forward1_train_256x448.json", rep)
    assert rep["train"]["mc_loss_rel"] <= 1e-4 and rep["train"]["bpp_mv_rel"] <= 1e-4, rep["train"]
    assert rep["train"]["grad_feature"] <= tol, rep["train"]
    assert stock_t[1] and all(v <= tol for v in rep["train"]["grad_params"].values()), rep["train"]
    tf = rep["train_full_forward"]
    assert tf["mse_loss_rel"] <= 1e-4 and tf["bpp_rel"] <= 1e-4, tf
    assert tf["grad_feature"] <= tol_f and tf["grad_params_max"] <= tol_f and any(k.startswith("res_codec") for k in stock_f[1]), tf


def test_compress_decompress_roundtrip_patched_and_streams_vs_stock(reference_modules):
    """``DeepSVC.compress`` -> ``decompress`` (``video_model.py:137-167``,
    ``image_model.py:201-302``) through the drop-in symbol pipeline and the C++ range coder;
    the stock run uses the oracle's pure-Python coder (256x448: seconds)."""
    import deepsvc_b200 as d
    modules, image_model, video_model = reference_modules
    H, W = 256, 448
    model = _make_model(video_model)
    ref, cur, sm, fea = _make_inputs(H, W)
    model.update(force=True)                     # shim tables (CPU python quantiser)
    with torch.no_grad():
        mv_s, res_s = model.compress(ref, cur, sm, fea)
        fea_s, rec_s, warped_s, pred_s = model.decompress(ref, mv_s, res_s, sm, fea)
    try:
        d.patch_reference(modules, video_model, image_model)
        assert d.swap_entropy_models(model) == 4
        model.update(force=True)                 # tables through the C++ quantiser
        with torch.no_grad():
            mv_p, res_p = model.compress(ref, cur, sm, fea)
            fea_p, rec_p, warped_p, pred_p = model.decompress(ref, mv_p, res_p, sm, fea)
            # encoder-side reconstruction of the same frame (eval-mode forward = round about the mean)
            fwd = model(ref, cur, sm, fea)
    finally:
        d.unpatch_reference()
    nbytes = lambda o: sum(len(s) for group in o["strings"] for s in group)  # noqa: E731
    rep = {"bytes_stock": nbytes(mv_s) + nbytes(res_s), "bytes_patched": nbytes(mv_p) + nbytes(res_p),
           "mv_y_identical": mv_s["strings"][0][0] == mv_p["strings"][0][0],
           "mv_z_identical": list(mv_s["strings"][1]) == list(mv_p["strings"][1]),
           "res_y_identical": res_s["strings"][0][0] == res_p["strings"][0][0],
           "res_z_identical": list(res_s["strings"][1]) == list(res_p["strings"][1]),
           "decoded_vs_stock_max_abs": (rec_p - rec_s).abs().max().item(),
           "decoded_vs_forward_max_abs": (rec_p - fwd[0].clamp(0, 1)).abs().max().item()}
    _dump("dropin_parity_codec_256x448.json", rep)
    # the decoder reproduces the encoder's reconstruction (decode(encode(x)) == forward's x_hat path)
    assert rep["decoded_vs_forward_max_abs"] <= 1e-3
    assert rep["decoded_vs_stock_max_abs"] <= 1e-3
    # same symbols and tables -> same bytes (the mv stream is produced before any drop-in output
    # other than the SpyNet warps reaches a conv; see the forward test for the tie-flip allowance)
    assert rep["mv_z_identical"] and rep["mv_y_identical"]
    assert abs(rep["bytes_patched"] - rep["bytes_stock"]) <= 0.001 * rep["bytes_stock"] + 8


def test_pframe_hotpath_vs_gpu_oracle_1080p(oracle):
    """The bound launch sequence at BASELINE configs[1] size (1, 1088, 1920) against the oracle
    run with stock torch on the same GPU (the reference's CUDA branch, modules.py:44-62)."""
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from deepsvc_b200.hotpath import PFrameHotPath
    dev = _dev()
    cpu_in = synthetic.make_pframe_inputs(B=1, H=1088, W=1920, seed=16)
    gpu_in = synthetic.to_device(cpu_in, dev)
    models_o, models = {}, {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb_o, gc_o = oracle.make_entropy_models(ch, seed=ch)
        eb = d.EntropyBottleneck(ch)
        eb.load_state_dict(eb_o.state_dict(), strict=False)
        models_o[name] = (eb_o.to(dev).eval(), gc_o.to(dev).eval())
        models[name] = (eb.to(dev).eval(), d.GaussianConditional(None).to(dev).eval())
    with torch.no_grad():
        want = oracle.pframe_hotpath(gpu_in, models_o)
    hp = PFrameHotPath(gpu_in, models)
    hp.capture()
    hp.replay()
    torch.cuda.synchronize()
    got = hp.results()
    for a, b in zip(got["spynet"] + [got["warped_frame"], got["warped_feature"]],
                    want["spynet"] + [want["warped_frame"], want["warped_feature"]]):
        err = (a - b).abs().max().item()
        assert err <= 1e-5 * max(1.0, b.abs().max().item()), (tuple(a.shape), err)
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6)
    for name in ("mv", "res"):
        assert torch.equal(got[f"{name}_y_hat"], want[f"{name}_y_hat"])
        assert torch.equal(got[f"{name}_z_hat"], want[f"{name}_z_hat"])
        assert _rel(got[f"bpp_{name}"], want[f"bpp_{name}"]) <= 1e-4
    assert _rel(got["bpp"], want["bpp"]) <= 1e-4


def test_iframe_codec_and_any_compressai_model_through_the_dropins(reference_modules):
    """the I-frame codec ``ICIP2020ResB`` (``image_model.py:440-488``: M = 320 in 10
    slices of 32 channels, N = 192) uses the same entropy ops; ``swap_entropy_models`` wires any
    model built on them.  Unmodified reference class, stock vs patched on the GPU: bit-identical."""
    import deepsvc_b200 as d
    modules, image_model, video_model = reference_modules
    torch.manual_seed(5)
    net = image_model.ICIP2020ResB().to(_dev()).eval()
    g = torch.Generator().manual_seed(6)
    x = torch.rand(1, 3, 256, 448, generator=g).to(_dev())
    with torch.no_grad():
        stock = net(x)
    try:
        d.patch_reference(modules, video_model, image_model)
        assert d.swap_entropy_models(net) == 2
        with torch.no_grad():
            patched = net(x)
    finally:
        d.unpatch_reference()
    assert torch.equal(patched["x_hat"], stock["x_hat"])
    for k in ("y", "z"):
        a, b = patched["likelihoods"][k], stock["likelihoods"][k]
        assert torch.allclose(a, b, rtol=2e-4, atol=1e-12)
        assert _rel(torch.log(a.double()).sum().item(), torch.log(b.double()).sum().item()) <= 1e-4
