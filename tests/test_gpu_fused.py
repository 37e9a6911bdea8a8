"""the fused few-channel warps against the oracle's restatement of the reference
statements they replace (modules.py:163-168, video_model.py:37-38)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("where", ["cpu", "cuda"])
@pytest.mark.parametrize("shape", [(1, 3, 32, 56), (2, 3, 136, 240), (1, 3, 2, 2), (1, 1, 34, 66), (1, 3, 1088, 1920)])
def test_spynet_level_warp(oracle, shape, where):
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    if where == "cpu" and H * W > 300 * 300:
        pytest.skip("full size: GPU oracle only")
    g = torch.Generator().manual_seed(H + W)
    im2 = torch.rand(B, C, H, W, generator=g)
    flow = synthetic.smooth_flow(B, max(H // 2, 1), max(W // 2, 1), g, sigma=3.0) if H >= 32 else \
        torch.randn(B, 2, H // 2, W // 2, generator=g)
    odev = _dev() if where == "cuda" else torch.device("cpu")
    d.set_flow_arithmetic(where)
    try:
        fu_ref, w_ref = oracle.spynet_level(im2.to(odev), flow.to(odev))
        fu, w = d.spynet_level_warp(im2.to(_dev()), flow.to(_dev()))
        # the unfused drop-in on the fused kernel's own flow_up is bit-identical
        assert torch.equal(d.torch_warp(im2.to(_dev()), fu), w)
    finally:
        d.set_flow_arithmetic("cuda")
    fu_ref, w_ref = fu_ref.to(_dev()), w_ref.to(_dev())
    assert (fu - fu_ref).abs().max().item() <= 1e-6 * max(1.0, fu_ref.abs().max().item())
    assert (w - w_ref).abs().max().item() <= 1e-5 * max(1.0, w_ref.abs().max().item())


@pytest.mark.parametrize("kind", ["smooth", "border"])
@pytest.mark.parametrize("shape", [(1, 3, 64, 96), (2, 3, 100, 132), (1, 3, 1088, 1920)])
def test_warp_with_mse(oracle, shape, kind):
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H * 3 + W)
    ref = torch.rand(B, C, H, W, generator=g).to(_dev())
    cur = torch.rand(B, C, H, W, generator=g).to(_dev())
    flow = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    w_ref, l_ref = oracle.warp_and_loss(ref, flow, cur)
    w, l = d.warp_with_mse(ref, flow, cur)
    assert torch.equal(w, d.torch_warp(ref, flow))
    assert (w - w_ref).abs().max().item() <= 1e-5
    assert abs(float(l) - float(l_ref)) <= 1e-5 * float(l_ref)
    assert float(d.warp_with_mse(ref, flow, cur)[1]) == float(l)  # fixed-order sum


def test_fused_ops_refuse_cpu_and_wide_tensors():
    import deepsvc_b200 as d
    x = torch.rand(1, 3, 8, 8)
    with pytest.raises(RuntimeError):
        d.spynet_level_warp(x, torch.zeros(1, 2, 4, 4))
    with pytest.raises(RuntimeError):
        d.spynet_level_warp(torch.rand(1, 8, 8, 8, device=_dev()), torch.zeros(1, 2, 4, 4, device=_dev()))


def test_fusions_are_differentiable_like_the_unfused_chains(oracle):
    """Gradients of spynet_level_warp (modules.py:163-168), warp_with_mse (video_model.py:37-38) and
    mc_blend (modules.py:436) against autograd of the reference's own statements on the GPU."""
    import deepsvc_b200 as d
    import torch.nn.functional as F
    dev = _dev()
    g = torch.Generator().manual_seed(21)

    def close(a, b, what):
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{what}: {err}"

    # SpyNet level: flow_up feeds the warp AND is used downstream (flow = flow_up + conv(...))
    im = torch.rand(2, 3, 48, 80, generator=g).to(dev)
    fl = (torch.randn(2, 2, 24, 40, generator=g) * 2).to(dev)
    c_up, c_w = torch.randn(2, 2, 48, 80, generator=g).to(dev), torch.randn(2, 3, 48, 80, generator=g).to(dev)
    res = []
    for fused in (False, True):
        a, b = im.clone().requires_grad_(True), fl.clone().requires_grad_(True)
        if fused:
            up, w = d.spynet_level_warp(a, b)
        else:
            up = F.interpolate(b, (48, 80), mode="bilinear", align_corners=False) * 2.0
            w = oracle.torch_warp(a, up)
        ((up * c_up).sum() + (w * c_w).sum()).backward()
        res.append((a.grad, b.grad))
    close(res[1][0], res[0][0], "spynet grad_im2")
    close(res[1][1], res[0][1], "spynet grad_flow")
    # warp + warp_loss
    ref = torch.rand(1, 3, 64, 96, generator=g).to(dev)
    cur = torch.rand(1, 3, 64, 96, generator=g).to(dev)
    mv = (torch.randn(1, 2, 64, 96, generator=g) * 3).to(dev)
    cot = torch.randn(1, 3, 64, 96, generator=g).to(dev)
    res = []
    for fused in (False, True):
        a, b, c = (t.clone().requires_grad_(True) for t in (ref, mv, cur))
        if fused:
            w, loss = d.warp_with_mse(a, b, c)
        else:
            w = oracle.torch_warp(a, b)
            loss = torch.mean((w - c).pow(2))
        (loss * 1000.0 + (w * cot).sum()).backward()
        res.append((a.grad, b.grad, c.grad))
    for x, y, nm in zip(res[1], res[0], ("grad_ref", "grad_flow", "grad_cur")):
        close(x, y, "warp_with_mse " + nm)
    # blend
    ws, wa, pr = (torch.rand(1, 3, 40, 56, generator=g).to(dev) for _ in range(3))
    res = []
    for fused in (False, True):
        a, b, c = (t.clone().requires_grad_(True) for t in (ws, wa, pr))
        out = d.mc_blend(a, b, c) if fused else a * b + (1 - a) * c
        (out * cot[:, :, :40, :56]).sum().backward()
        res.append((a.grad, b.grad, c.grad))
    for x, y in zip(res[1], res[0]):
        close(x, y, "mc_blend")


@pytest.mark.parametrize("shape", [(1, 3, 64, 96), (2, 3, 33, 47), (1, 3, 1088, 1920), (1, 1, 1, 3)])
def test_mc_blend_bit_identical(shape):
    """modules.py:436: w * warped + (1 - w) * pred, bit-identical to the torch expression."""
    import deepsvc_b200 as d
    g = torch.Generator().manual_seed(sum(shape))
    w = torch.rand(shape, generator=g).to(_dev())
    a = torch.randn(shape, generator=g).to(_dev())
    b = torch.randn(shape, generator=g).to(_dev())
    assert torch.equal(d.mc_blend(w, a, b), w * a + (1 - w) * b)
    # an unaligned view takes the scalar path
    wv, av, bv = (t.reshape(-1)[1:].contiguous()[1:] for t in (w, a, b))
    wv, av, bv = (t.reshape(-1)[1:] for t in (w, a, b))
    assert torch.equal(d.mc_blend(wv, av, bv), wv * av + (1 - wv) * bv)


def test_lrp_add_bit_identical_and_gradients():
    """image_model.py:185-188 (`lrp = 0.5 * torch.tanh(lrp); y_hat_slice += lrp`) in one launch."""
    import deepsvc_b200 as d
    g = torch.Generator().manual_seed(8)
    dev = torch.device("cuda:0")
    for shape in ((1, 8, 68, 120), (8, 12, 16, 16), (1, 3, 5, 7)):
        y = (torch.randn(shape, generator=g) * 3).to(dev)
        l = (torch.randn(shape, generator=g) * 2).to(dev)
        want = y + 0.5 * torch.tanh(l)
        assert torch.equal(d.lrp_add(y, l), want)
        y2 = y.clone()
        assert d.lrp_add(y2, l, inplace=True) is y2 and torch.equal(y2, want)
        # a batch-strided slice (y.chunk(8, 1)[i] with B > 1) is copied once
        big = (torch.randn(shape[0], shape[1] * 2, *shape[2:], generator=g)).to(dev)
        ys = big.chunk(2, 1)[1]
        assert torch.equal(d.lrp_add(ys, l), ys + 0.5 * torch.tanh(l))
        ya, la = y.clone().requires_grad_(True), l.clone().requires_grad_(True)
        yb, lb = y.clone().requires_grad_(True), l.clone().requires_grad_(True)
        cot = torch.randn(shape, generator=g).to(dev)
        d.lrp_add(ya, la).backward(cot)
        (yb + 0.5 * torch.tanh(lb)).backward(cot)
        assert torch.equal(ya.grad, yb.grad) and torch.allclose(la.grad, lb.grad, rtol=1e-5, atol=1e-7)
