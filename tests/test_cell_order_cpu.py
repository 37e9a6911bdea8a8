"""The identity behind csrc/warp_bwd_cell.cu, on the CPU: gradients evaluated cell by cell
(oracle/cell_order.py) equal autograd of the reference's torch_warp (modules.py:25-62 ->
grid_sampler_2d_backward) -- float64, so the comparison is to rounding."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("kind", ["smooth", "stress", "border", "gentle", "integer"])
@pytest.mark.parametrize("shape", [(1, 3, 16, 24), (2, 5, 33, 47), (1, 2, 7, 64)])
def test_cell_order_equals_autograd_of_the_reference_warp(oracle, shape, kind):
    from oracle.cell_order import cell_order_backward
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) + len(kind))
    inp = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    if kind == "integer":   # coordinates exactly on cell corners (weights 0 / 1, clamped borders)
        flow = torch.randint(-6, 7, (B, 2, H, W), generator=g).double()
    else:
        flow = synthetic.make_flow(kind, B, H, W, g).double()
    gout = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    a, f = inp.clone().requires_grad_(True), flow.clone().requires_grad_(True)
    oracle._grid_cache.clear()   # the cache is keyed on the flow's size, not its dtype
    oracle.torch_warp(a, f).backward(gout)
    oracle._grid_cache.clear()
    gin, gflow = cell_order_backward(gout.numpy(), inp.numpy(), flow.numpy())
    assert np.abs(gin - a.grad.numpy()).max() <= 1e-10 * max(1.0, float(a.grad.abs().max()))
    assert np.abs(gflow - f.grad.numpy()).max() <= 1e-9 * max(1.0, float(f.grad.abs().max()))
    # every element of grad_input is produced (no zero-fill needed): the plane sums are the taps' weights
    ones, _ = cell_order_backward(np.ones_like(gout.numpy()), inp.numpy(), flow.numpy())
    assert np.allclose(ones.sum((2, 3)), H * W, rtol=1e-12)
