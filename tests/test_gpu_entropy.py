"""GPU parity of the fused entropy kernels (through the C ABI) against the CPU oracle
(restatement of compressai 1.2.1) and the golden fixtures.

Bars (north_star): quantised latents / symbols / indexes bit-exact; bpp within 1e-4
relative.  Likelihood tensors: elementwise rtol 2e-4 (the value is a difference of two
erfc's; CUDA erfcf and the CPU libm differ in the last ulps, amplified by cancellation
when scale >> 1) with atol 1e-9 at the likelihood floor."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _models(oracle, ch, seed=0):
    import deepsvc_b200 as d
    eb_o, gc_o = oracle.make_entropy_models(ch, seed=seed)
    eb = d.EntropyBottleneck(ch)
    eb.load_state_dict(eb_o.state_dict(), strict=False)
    gc = d.GaussianConditional(None)
    gc.scale_table = gc_o.scale_table.clone()
    return (eb_o.eval(), gc_o.eval()), (eb.to(_dev()).eval(), gc.to(_dev()).eval())


def _bits(lik):
    return float(torch.log(lik.double()).sum() / -math.log(2))


def test_gc_known_answers(oracle, golden_dir):
    d = np.load(os.path.join(golden_dir, "entropy_kat.npz"))
    _, (eb, gc) = _models(oracle, 8, seed=3)
    x, s, m = (torch.from_numpy(d[k]).to(_dev()) for k in ("x", "scales", "means"))
    out, lik = gc(x, s, m)
    assert torch.equal(out.cpu(), torch.from_numpy(d["out"]))
    np.testing.assert_allclose(lik.cpu().numpy(), d["lik"], rtol=2e-4, atol=1e-9)
    assert torch.equal(gc.quantize(x, "symbols", m).cpu(), torch.from_numpy(d["symbols"]))
    assert torch.equal(gc.quantize(x, "dequantize", m).cpu(), torch.from_numpy(d["out"]))
    assert torch.equal(gc.build_indexes(s).cpu(), torch.from_numpy(d["indexes"]))
    tbl = torch.from_numpy(d["table"]).to(_dev()).reshape(1, 1, 1, -1)
    assert torch.equal(gc.build_indexes(tbl).cpu(), torch.from_numpy(d["table_indexes"]))
    sym, idx, y_hat = gc.quantize_and_index(x, s, m)
    assert torch.equal(sym.cpu(), torch.from_numpy(d["symbols"]))
    assert torch.equal(idx.cpu(), torch.from_numpy(d["indexes"]))
    assert torch.equal(y_hat.cpu(), torch.from_numpy(d["out"]))
    assert torch.equal(gc.dequantize(sym, m).cpu(), torch.from_numpy(d["out"]))


def test_callsite_golden(oracle, golden_dir):
    """Tensors recorded at the reference's own call sites inside DeepSVC.forward."""
    import deepsvc_b200 as dsvc
    d = np.load(os.path.join(golden_dir, "callsite_deepsvc_64x128.npz"))
    pixels = 64 * 128
    total_bits = 0.0
    for cname, ch in (("mv", 64), ("res", 96)):
        eb = dsvc.EntropyBottleneck(ch)
        with torch.no_grad():
            for n, p in eb.named_parameters():
                p.copy_(torch.from_numpy(d[f"eb_{cname}_param_{n}"]))
        eb = eb.to(_dev()).eval()
        z = torch.from_numpy(d[f"eb_{cname}_z"]).to(_dev())
        with torch.no_grad():
            z_out, z_lik = eb(z)
            z_hat, part = eb.likelihood_bits(z)
        assert torch.equal(z_out.cpu(), torch.from_numpy(d[f"eb_{cname}_out"]))
        assert torch.equal(z_hat.cpu(), torch.from_numpy(d[f"eb_{cname}_out"]))
        np.testing.assert_allclose(z_lik.cpu().numpy(), d[f"eb_{cname}_lik"], rtol=2e-4, atol=1e-9)
        total_bits += float(part.sum()) / -math.log(2)
        gc = dsvc.GaussianConditional(None).to(_dev()).eval()
        for k in range(8):
            x, s, m = (torch.from_numpy(d[f"gc_{cname}_{key}"][k]).to(_dev()) for key in ("x", "scales", "means"))
            out, lik = gc(x, s, m)
            assert torch.equal(out.cpu(), torch.from_numpy(d[f"gc_{cname}_out"][k]))
            np.testing.assert_allclose(lik.cpu().numpy(), d[f"gc_{cname}_lik"][k], rtol=2e-4, atol=1e-9)
            y_hat, part = gc.likelihood_bits(x, s, m)
            assert torch.equal(y_hat.cpu(), torch.from_numpy(d[f"gc_{cname}_out"][k]))
            total_bits += float(part.sum()) / -math.log(2)
    ref_bpp = float(d["bpp"])
    assert abs(total_bits / pixels - ref_bpp) <= 1e-4 * ref_bpp


@pytest.mark.parametrize("shape", [(1, 8, 16, 28), (1, 12, 68, 120), (8, 8, 16, 16), (2, 12, 5, 7), (1, 1, 1, 1), (3, 5, 3, 11)])
@pytest.mark.parametrize("training", [False, True])
def test_gc_forward_vs_oracle(oracle, shape, training):
    from deepsvc_b200 import synthetic
    (_, gc_o), (_, gc) = _models(oracle, 8)
    g = torch.Generator().manual_seed(sum(shape))
    B, C, h, w = shape
    y, s, m = synthetic.make_latents(B, C, h, w, g, tie_frac=0.05, tail_frac=0.01)
    noise = torch.rand(y.shape, generator=g) - 0.5
    if training:
        ref_out = y + noise
        ref_lik = gc_o.likelihood_lower_bound(gc_o._likelihood(ref_out, s, m))
    else:
        ref_out, ref_lik = gc_o(y, s, m)
    ref_yhat = oracle.ste_round(y - m) + m
    yd, sd, md = y.to(_dev()), s.to(_dev()), m.to(_dev())
    y_hat, lik, part = gc.forward_fused(yd, sd, md, training=training, noise=noise.to(_dev()) if training else None)
    assert torch.equal(y_hat.cpu(), ref_yhat)
    np.testing.assert_allclose(lik.cpu().numpy(), ref_lik.numpy(), rtol=2e-4, atol=1e-9)
    rb = _bits(ref_lik)
    assert abs(float(part.sum()) / -math.log(2) - rb) <= 1e-4 * abs(rb)
    if not training:
        out, lik2 = gc(yd, sd, md)
        assert torch.equal(out.cpu(), ref_out)
        assert torch.equal(lik2, lik)
        assert torch.equal(gc.quantize(yd, "symbols", md).cpu(), gc_o.quantize(y, "symbols", m))
        assert torch.equal(gc.build_indexes(sd).cpu(), gc_o.build_indexes(s))


def test_gc_batch_strided_slices(oracle):
    """y.chunk(8, 1) slices of a batched latent are not contiguous (image_model.py:164)."""
    from deepsvc_b200 import synthetic
    (_, gc_o), (_, gc) = _models(oracle, 8)
    g = torch.Generator().manual_seed(4)
    y, s, m = synthetic.make_latents(4, 64, 6, 10, g)
    yd, sd, md = y.to(_dev()), s.to(_dev()), m.to(_dev())
    for ys, ss, ms, yo, so, mo in zip(yd.chunk(8, 1), sd.chunk(8, 1), md.chunk(8, 1),
                                      y.chunk(8, 1), s.chunk(8, 1), m.chunk(8, 1)):
        assert not ys.is_contiguous()
        out, lik = gc(ys, ss, ms)
        ro, rl = gc_o(yo, so, mo)
        assert torch.equal(out.cpu(), ro)
        np.testing.assert_allclose(lik.cpu().numpy(), rl.numpy(), rtol=2e-4, atol=1e-9)
        assert torch.equal(gc.build_indexes(ss).cpu(), gc_o.build_indexes(so))
    # a view the kernels cannot read in place (e.g. the crops of image_model.py:171,175 on an
    # unpadded input) is copied once, like the stock eager ops; strict mode raises instead
    from deepsvc_b200 import entropy as E
    out, lik = gc(yd[:, :, :, ::2], sd[:, :, :, ::2], md[:, :, :, ::2])
    ro, rl = gc_o(y[:, :, :, ::2], s[:, :, :, ::2], m[:, :, :, ::2])
    assert torch.equal(out.cpu(), ro)
    np.testing.assert_allclose(lik.cpu().numpy(), rl.numpy(), rtol=2e-4, atol=1e-9)
    E.STRICT_STRIDES = True
    try:
        with pytest.raises(RuntimeError, match="strides"):
            gc(yd[:, :, :, ::2], sd[:, :, :, ::2], md[:, :, :, ::2])
    finally:
        E.STRICT_STRIDES = False


def test_gc_means_none_and_empty(oracle):
    (_, gc_o), (_, gc) = _models(oracle, 8)
    g = torch.Generator().manual_seed(1)
    y = torch.randn(1, 4, 3, 5, generator=g) * 3
    s = torch.rand(1, 4, 3, 5, generator=g) + 0.05
    out, lik = gc(y.to(_dev()), s.to(_dev()))
    ro, rl = gc_o(y, s)
    assert torch.equal(out.cpu(), ro)
    np.testing.assert_allclose(lik.cpu().numpy(), rl.numpy(), rtol=2e-4, atol=1e-9)
    e = torch.empty(0, 4, 3, 5, device=_dev())
    out, lik = gc(e, e, e)
    assert out.shape == e.shape and lik.shape == e.shape


@pytest.mark.parametrize("shape", [(1, 64, 4, 7), (1, 96, 17, 30), (8, 64, 4, 4), (2, 5, 3, 3)])
@pytest.mark.parametrize("training", [False, True])
def test_eb_forward_vs_oracle(oracle, shape, training):
    B, C, h, w = shape
    (eb_o, _), (eb, _) = _models(oracle, C, seed=C)
    g = torch.Generator().manual_seed(sum(shape))
    z = torch.randn(B, C, h, w, generator=g) * 3
    noise = torch.rand(z.shape, generator=g) - 0.5
    med = eb_o._get_medians().detach()
    ref_zhat = oracle.ste_round(z - med) + med
    with torch.no_grad():
        if training:
            vals = (z + noise).permute(1, 0, 2, 3).reshape(C, 1, -1)
            ref_lik = eb_o.likelihood_lower_bound(eb_o._likelihood(vals)).reshape(C, B, h, w).permute(1, 0, 2, 3)
        else:
            ref_out, ref_lik = eb_o(z)
        z_hat, lik, part = eb.forward_fused(z.to(_dev()), training=training,
                                            noise=noise.to(_dev()) if training else None)
    assert torch.equal(z_hat.cpu(), ref_zhat)
    np.testing.assert_allclose(lik.cpu().numpy(), ref_lik.numpy(), rtol=2e-4, atol=1e-9)
    rb = _bits(ref_lik)
    assert abs(float(part.sum()) / -math.log(2) - rb) <= 1e-4 * abs(rb)
    if not training:
        with torch.no_grad():
            out, lik2 = eb(z.to(_dev()))
        assert torch.equal(out.cpu(), ref_out)


def test_gc_backward_vs_oracle_autograd(oracle):
    """Gradients of the likelihood (incl. both LowerBound pass-through rules) in noise
    mode, and of the straight-through y_hat; tolerance 1e-4 of the gradient's scale."""
    from deepsvc_b200 import synthetic
    (_, gc_o), (_, gc) = _models(oracle, 8)
    g = torch.Generator().manual_seed(21)
    y, s, m = synthetic.make_latents(2, 8, 9, 13, g, tie_frac=0.0, tail_frac=0.02)
    noise = torch.rand(y.shape, generator=g) - 0.5
    w_l = torch.randn(y.shape, generator=g)
    w_y = torch.randn(y.shape, generator=g)

    def run(fn, dev):
        yy, ss, mm = (t.clone().to(dev).requires_grad_(True) for t in (y, s, m))
        lik, y_hat = fn(yy, ss, mm, noise.to(dev))
        loss = (torch.log(lik) * w_l.to(dev)).sum() + (y_hat * w_y.to(dev)).sum()
        loss.backward()
        return [t.grad.cpu() for t in (yy, ss, mm)]

    def ref_fn(yy, ss, mm, nz):
        lik = gc_o.likelihood_lower_bound(gc_o._likelihood(yy + nz, ss, mm))
        return lik, oracle.ste_round(yy - mm) + mm

    def got_fn(yy, ss, mm, nz):
        y_hat, lik, _ = gc.forward_fused(yy, ss, mm, training=True, noise=nz)
        return lik, y_hat

    for a, b, nm in zip(run(got_fn, _dev()), run(ref_fn, "cpu"), ("g_y", "g_scale", "g_mean")):
        scale = max(1.0, b.abs().max().item())
        assert (a - b).abs().max().item() <= 1e-4 * scale, nm


def test_eb_backward_vs_oracle_autograd(oracle):
    C = 6
    (eb_o, _), (eb, _) = _models(oracle, C, seed=2)
    g = torch.Generator().manual_seed(22)
    z = torch.randn(3, C, 5, 4, generator=g) * 3
    noise = torch.rand(z.shape, generator=g) - 0.5
    w_l = torch.randn(z.shape, generator=g)

    zo = z.clone().requires_grad_(True)
    vals = (zo + noise).permute(1, 0, 2, 3).reshape(C, 1, -1)
    lik_o = eb_o.likelihood_lower_bound(eb_o._likelihood(vals)).reshape(C, 3, 5, 4).permute(1, 0, 2, 3)
    (torch.log(lik_o) * w_l).sum().backward()

    eb.train()
    zd = z.clone().to(_dev()).requires_grad_(True)
    _, lik, _ = eb.forward_fused(zd, training=True, noise=noise.to(_dev()))
    (torch.log(lik) * w_l.to(_dev())).sum().backward()
    assert (zd.grad.cpu() - zo.grad).abs().max().item() <= 1e-4 * max(1.0, zo.grad.abs().max().item())
    for (n, p), (n2, p2) in zip(eb.named_parameters(), eb_o.named_parameters()):
        assert n == n2
        if n == "quantiles":
            continue  # noise mode: medians unused, grad None / zero on both sides
        assert p.grad is not None, n
        scale = max(1.0, p2.grad.abs().max().item())
        assert (p.grad.cpu() - p2.grad).abs().max().item() <= 2e-4 * scale, n


def test_full_size_1080p_checksum(oracle):
    """Config 2 size: 12-ch slice at 68x120 -- sum of per-CTA partials equals the log-sum
    of the materialised likelihood tensor (a checksum of checksums), symbols consistent
    with y_hat."""
    from deepsvc_b200 import synthetic
    _, (_, gc) = _models(oracle, 8)
    g = torch.Generator().manual_seed(8)
    y, s, m = synthetic.make_latents(1, 12, 68, 120, g)
    yd, sd, md = y.to(_dev()), s.to(_dev()), m.to(_dev())
    y_hat, lik, part = gc.forward_fused(yd, sd, md, training=False)
    a = float(part.sum())
    b = float(torch.log(lik.double()).sum())
    assert abs(a - b) <= 1e-6 * abs(b)
    sym, idx, y_hat2 = gc.quantize_and_index(yd, sd, md)
    assert torch.equal(y_hat2, y_hat)
    assert torch.equal(sym.float() + md, y_hat)
    assert int(idx.min()) >= 0 and int(idx.max()) <= 63


@pytest.mark.parametrize("case", [("ICIP2020ResB", 1, 320, 192, 10, 16, 28), ("ICIP2020ResB", 1, 320, 192, 10, 68, 120),
                                  ("cFeatureCompress", 4, 72, 72, 6, 16, 16)])
def test_iframe_codec_shapes_f4(oracle, case):
    """the I-frame codec (ICIP2020ResB, image_model.py:440-488: M = 320 = 10 slices x 32
    channels, N = 192 hyper-latent channels) and the semantic layer's cFeatureCompress
    (semantic_layer.py:1189-1200,1324-1372: N = 72 = 6 slices x 12 channels, batch 4 of 256^2 crops ->
    y [4,72,16,16], z [4,72,4,4]) call the same ops in the same pattern.  y_hat / z_hat bit-exact, bits
    within 1e-4 relative, for the eval path and the fused bit-sum path."""
    import math
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import synthetic
    dev = torch.device("cuda:0")
    _, B, M, N, n_slices, h, w = case
    g = torch.Generator().manual_seed(M + h)
    y, scales, means = synthetic.make_latents(B, M, h, w, g)
    z = torch.randn(B, N, max(h // 4, 1), max(w // 4, 1), generator=g) * 3.0
    eb_o, gc_o = oracle.make_entropy_models(N, seed=N)
    eb_o.eval(), gc_o.eval()
    eb = dsvc.EntropyBottleneck(N)
    eb.load_state_dict(eb_o.state_dict(), strict=False)
    eb, gc = eb.to(dev).eval(), dsvc.GaussianConditional(None).to(dev).eval()
    bits_ref = bits_got = 0.0
    with torch.no_grad():
        _, zl_ref = eb_o(z)
        off = eb_o._get_medians()
        zh_ref = oracle.ste_round(z - off) + off
        zh, zl, zpart = eb.forward_fused(z.to(dev))
        assert torch.equal(zh.cpu(), zh_ref)
        bits_ref += torch.log(zl_ref).sum().item()
        bits_got += torch.log(zl).sum().item()
        fused = zpart.sum().item()
        for y_s, s_s, m_s in zip(y.chunk(n_slices, 1), scales.chunk(n_slices, 1), means.chunk(n_slices, 1)):
            assert y_s.shape[1] == M // n_slices
            _, lik_ref = gc_o(y_s, s_s, m_s)
            yh_ref = oracle.ste_round(y_s - m_s) + m_s
            yh, lik, part = gc.forward_fused(y_s.to(dev), s_s.to(dev), m_s.to(dev))
            assert torch.equal(yh.cpu(), yh_ref)
            bits_ref += torch.log(lik_ref).sum().item()
            bits_got += torch.log(lik).sum().item()
            fused += part.sum().item()
    assert abs(bits_got - bits_ref) <= 1e-4 * abs(bits_ref)
    assert abs(fused - bits_ref) <= 1e-4 * abs(bits_ref)
    assert math.isfinite(fused)


def test_bottleneck_parameter_packing_fused_vs_eager():
    """dsvc_eb_pack_f32 / _bwd_f32 (one launch each way) against the eager torch chain it replaces
    (softplus / tanh / cat of the 15 raw tensors) and its autograd."""
    import torch.nn.functional as F
    import deepsvc_b200 as d
    from deepsvc_b200.entropy import pack_bottleneck_params
    dev = _dev()
    g = torch.Generator().manual_seed(4)
    eb = d.EntropyBottleneck(7)
    with torch.no_grad():
        for p in eb.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 2)
        eb._matrix0[0, 0, 0] = 25.0       # softplus threshold branch
    eb = eb.to(dev)

    def eager():
        C, parts = eb.channels, []
        for i in range(5):
            parts.append(F.softplus(getattr(eb, f"_matrix{i}")).reshape(C, -1))
            parts.append(getattr(eb, f"_bias{i}").reshape(C, -1))
            if i < 4:
                parts.append(torch.tanh(getattr(eb, f"_factor{i}")).reshape(C, -1))
        parts.append(eb.quantiles[:, 0, 1:2])
        parts.append(torch.zeros(C, 1, device=dev))
        return torch.cat(parts, 1)

    cot = torch.randn(7, 60, generator=g).to(dev)
    want = eager()
    want.backward(cot)
    ref_grads = [p.grad.clone() for p in eb.parameters()]
    for p in eb.parameters():
        p.grad = None
    got = pack_bottleneck_params(eb)
    assert torch.equal(got, want.detach())
    got.backward(cot)
    for p, r in zip(eb.parameters(), ref_grads):
        assert torch.allclose(p.grad, r, rtol=1e-6, atol=1e-7)
