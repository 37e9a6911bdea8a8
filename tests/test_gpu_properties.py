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
