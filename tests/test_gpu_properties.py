"""Property tests on random shapes (hypothesis) against the oracle: the warp (forward and both
gradients, every kernel family picked by the dispatcher) and the Gaussian-conditional launch.
SURVEY.md 8c(3)."""
import math

import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

pytestmark = pytest.mark.gpu

_SET = dict(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def _dev():
    return torch.device("cuda:0")


@settings(**_SET)
@given(B=st.integers(1, 3), C=st.integers(1, 20), H=st.integers(2, 90), W=st.integers(2, 150),
       kind=st.sampled_from(["smooth", "stress", "border", "integer"]), seed=st.integers(0, 10_000))
def test_warp_forward_backward_random_shapes(B, C, H, W, kind, seed):
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from oracle import reference_ops as R
    g = torch.Generator().manual_seed(seed)
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    if kind == "integer":
        flow0 = torch.randint(-5, 6, (B, 2, H, W), generator=g).float().to(_dev())
    else:
        flow0 = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    res = []
    for fn in (R.torch_warp, d.torch_warp):
        inp = inp0.clone().requires_grad_(True)
        flow = flow0.clone().requires_grad_(True)
        out = fn(inp, flow)
        out.backward(gout)
        res.append((out.detach(), inp.grad, flow.grad))
    (o_r, gi_r, gf_r), (o, gi, gf) = res
    assert (o - o_r).abs().max().item() <= 1e-5 * max(1.0, o_r.abs().max().item())
    assert (gi - gi_r).abs().max().item() <= 1e-4 * max(1.0, gi_r.abs().max().item())
    assert (gf - gf_r).abs().max().item() <= 1e-4 * max(1.0, gf_r.abs().max().item())


@settings(**_SET)
@given(B=st.integers(1, 3), C=st.integers(1, 24), h=st.integers(1, 20), w=st.integers(1, 33),
       nslice=st.sampled_from([1, 2, 3, 8]), training=st.booleans(), seed=st.integers(0, 10_000))
def test_gaussian_conditional_random_shapes(B, C, h, w, nslice, training, seed):
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from oracle import reference_ops as R
    g = torch.Generator().manual_seed(seed)
    y, s, m = synthetic.make_latents(B, C, h, w, g)
    noise = torch.rand(y.shape, generator=g) - 0.5
    gc_o = R.GaussianConditional(None)
    gc = d.GaussianConditional(None).to(_dev())
    gc_o.train(training), gc.train(training)
    yd, sd, md, nd = (t.to(_dev()) for t in (y, s, m, noise))
    bits_ref = bits = 0.0
    for ys, ss, ms, ns, yo, so, mo, no in zip(*(t.chunk(nslice, 1) for t in (yd, sd, md, nd, y, s, m, noise))):
        if training:
            lik_ref = gc_o.likelihood_lower_bound(gc_o._likelihood(yo + no, so, mo))
        else:
            _, lik_ref = gc_o(yo, so, mo)
        y_hat, lik, part = gc.forward_fused(ys, ss, ms, noise=ns if training else None)
        assert torch.equal(y_hat.cpu(), R.ste_round(yo - mo) + mo)           # bit-exact (batch-strided slices)
        assert torch.allclose(lik.cpu(), lik_ref, rtol=2e-4, atol=1e-9)
        bits_ref += torch.log(lik_ref.double()).sum().item()
        bits += part.sum().item()
    assert abs(bits - bits_ref) <= 1e-4 * max(abs(bits_ref), 1e-6)
    assert math.isfinite(bits)


@settings(**dict(_SET, max_examples=30))
@given(B=st.integers(1, 2), C=st.integers(1, 12), Hq=st.integers(1, 24), Wq=st.integers(1, 40),
       kind=st.sampled_from(["smooth", "stress", "border", "integer", "collapse"]), need_flow=st.booleans(),
       seed=st.integers(0, 10_000))
def test_warp_backward_kernels_agree_random_shapes(B, C, Hq, Wq, kind, need_flow, seed):
    """The four backward kernels -- per-pixel scatter, shared-memory staged scatter, destination-owned
    gather + fix-up launch, cell-order -- on random shapes (W a multiple of 4: the staged kernels' eligibility) and
    flow families, incl. flows that collapse many pixels onto one source element (list overflow, flagged
    tiles): same gradients up to the fp32 summation order, grad_input needs no zero-fill."""
    from deepsvc_b200 import _lib, synthetic, warp as warp_mod
    from deepsvc_b200.warp import warp_backward
    H, W = 4 * Hq, 4 * Wq
    g = torch.Generator().manual_seed(seed)
    inp = torch.randn(B, C, H, W, generator=g).to(_dev())
    if kind == "integer":
        flow = torch.randint(-5, 6, (B, 2, H, W), generator=g).float().to(_dev())
    elif kind == "collapse":
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
        flow = torch.cat([(W / 2 + 0.3) - xs, torch.randn(B, 1, H, W, generator=g)], 1).contiguous().to(_dev())
    else:
        flow = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    lib = _lib.load()
    res = {}
    for name, algo in (("direct", _lib.WARP_BWD_DIRECT), ("staged", _lib.WARP_BWD_STAGED), ("gather", _lib.WARP_BWD_GATHER),
                       ("cell", _lib.WARP_BWD_CELL)):
        _lib.check(lib.dsvc_set_warp_bwd_algo(algo), "algo")
        min_c = warp_mod.CELL_MIN_CHANNELS
        warp_mod.CELL_MIN_CHANNELS = 1   # the cell-order kernel takes any C once its workspace is there
        try:
            res[name] = warp_backward(gout, inp, flow, True, need_flow)
        finally:
            warp_mod.CELL_MIN_CHANNELS = min_c
            lib.dsvc_set_warp_bwd_algo(_lib.WARP_BWD_AUTO)
    for name in ("staged", "gather", "cell"):
        for a, b, nm in zip(res[name], res["direct"], ("grad_input", "grad_flow")):
            if b is None:
                assert a is None
                continue
            err = (a - b).abs().max().item()
            assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{name} {nm} {err}"
