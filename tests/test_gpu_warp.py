"""GPU parity of the warp kernels (through the C ABI) against the CPU oracle
(= the reference's torch_warp arithmetic), the golden fixtures, and stock torch CUDA.

Tolerance (north_star): warped tensors within 1e-5 relative in fp32.  Stated here as
max|a-b| <= 1e-5 * max(1, max|ref|) plus an elementwise allclose(rtol=1e-5, atol=2e-6)
(atol covers cancellation in the 4-tap sum; observed error is ~3e-7).
"""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ALGOS = ["gather", "auto"]


def _dev():
    return torch.device("cuda:0")


def assert_warp_close(got, ref, what=""):
    got = got.detach().cpu()
    ref = ref.detach().cpu()
    err = (got - ref).abs().max().item()
    tol = 1e-5 * max(1.0, ref.abs().max().item())
    assert err <= tol, f"{what}: max abs err {err:.3e} > {tol:.3e}"
    assert torch.allclose(got, ref, rtol=1e-5, atol=2e-6), what


def _warp(inp, flow, mode, algo):
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib
    fm = {"cpu": _lib.FLOW_TRUE_DIVIDE, "cuda": _lib.FLOW_MUL_RECIPROCAL}[mode]
    al = {"gather": _lib.WARP_GATHER, "auto": _lib.WARP_AUTO, "tma": _lib.WARP_TMA}[algo]
    return d.warp_forward(inp, flow, flow_mode=fm, algo=al)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "warp_*.npz"))))
def test_golden_fixtures(path, algo):
    """Fixtures produced by the reference's own modules.torch_warp (CPU branch)."""
    d = np.load(path)
    inp, flow = torch.from_numpy(d["input"]).to(_dev()), torch.from_numpy(d["flow"]).to(_dev())
    assert_warp_close(_warp(inp, flow, "cpu", algo), torch.from_numpy(d["out"]), os.path.basename(path))


SHAPES = [  # (B, C, H, W): the six per-frame shapes of config 1 + ragged / batched cases
    (1, 3, 32, 56), (1, 3, 64, 112), (1, 3, 128, 224), (1, 3, 256, 448), (1, 64, 256, 448),
    (8, 3, 32, 32), (8, 64, 64, 64), (2, 5, 17, 23), (1, 1, 2, 7), (1, 2, 9, 2), (3, 7, 33, 130),  # (H or W == 1 is NaN in the reference itself: 0/0)
]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("kind", ["smooth", "stress", "border"])
@pytest.mark.parametrize("shape", SHAPES)
def test_vs_cpu_oracle(oracle, shape, kind, algo):
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) * 7 + len(kind))
    inp = torch.randn(B, C, H, W, generator=g)
    if kind == "smooth":
        flow = synthetic.smooth_flow(B, H, W, g)
    elif kind == "stress":
        flow = synthetic.stress_flow(B, H, W, g)
    else:
        flow = synthetic.border_flow(B, H, W, g, margin=max(1, min(H, W) // 4), reach=40.0)
    ref = oracle.torch_warp(inp, flow)
    got = _warp(inp.to(_dev()), flow.to(_dev()), "cpu", algo)
    assert_warp_close(got, ref, f"{shape} {kind} {algo}")


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("shape", [(1, 3, 256, 448), (2, 64, 128, 192), (1, 5, 31, 47)])
def test_vs_stock_torch_cuda(oracle, shape, algo):
    """The reference's CUDA branch (modules.py:44-62) run with stock torch on this GPU."""
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(7)
    inp = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow = synthetic.smooth_flow(B, H, W, g, sigma=8.0).to(_dev())
    ref = oracle.torch_warp(inp, flow)  # same restatement, tensors on the GPU
    got = _warp(inp, flow, "cuda", algo)
    assert_warp_close(got, ref, f"{shape} cuda-branch {algo}")


@pytest.mark.parametrize("algo", ALGOS)
def test_exact_properties(algo):
    g = torch.Generator().manual_seed(3)
    inp = torch.randn(2, 6, 40, 72, generator=g).to(_dev())
    zero = torch.zeros(2, 2, 40, 72, device=_dev())
    out = _warp(inp, zero, "cuda", algo)
    assert (out - inp).abs().max().item() <= 2e-4  # identity up to (g+1)/2*(W-1) rounding
    # linearity in the input: warp(a*x + y) == a*warp(x) + warp(y) within fp32 rounding
    flow = (torch.randn(2, 2, 40, 72, generator=g) * 3).to(_dev())
    x, y = inp, torch.randn(2, 6, 40, 72, generator=g).to(_dev())
    lhs = _warp(2.0 * x + y, flow, "cuda", algo)
    rhs = 2.0 * _warp(x, flow, "cuda", algo) + _warp(y, flow, "cuda", algo)
    assert torch.allclose(lhs, rhs, rtol=1e-5, atol=1e-5)
    # a constant image stays constant under any flow (weights sum to 1)
    c = torch.full((1, 3, 24, 40), 0.75, device=_dev())
    fl = (torch.randn(1, 2, 24, 40, generator=g) * 50).to(_dev())
    assert (_warp(c, fl, "cuda", algo) - 0.75).abs().max().item() <= 1e-6
    # far outside -> border value
    fl = torch.zeros(1, 2, 24, 40, device=_dev())
    fl[:, 0] = 1e4
    img = torch.randn(1, 3, 24, 40, generator=g).to(_dev())
    out = _warp(img, fl, "cuda", algo)
    assert (out - img[:, :, :, -1:].expand(-1, -1, -1, 40)).abs().max().item() <= 1e-4


@pytest.mark.parametrize("kind", ["smooth", "stress", "border"])
@pytest.mark.parametrize("shape", [(2, 5, 40, 72), (1, 3, 32, 64), (1, 16, 100, 132), (3, 7, 33, 128),
                                   (1, 8, 8, 4), (2, 9, 65, 260)])
def test_tma_forced(oracle, shape, kind):
    """The TMA-staged kernel on shapes the auto policy would not pick: odd channel
    counts (partial last channel group), batch > 1 (plane index crosses items), partial
    edge tiles, and wild flows (per-tile in-kernel fallback)."""
    from deepsvc_b200 import synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape) * 3 + len(kind))
    inp = torch.randn(B, C, H, W, generator=g)
    if kind == "smooth":
        flow = synthetic.smooth_flow(B, H, W, g)
    elif kind == "stress":
        flow = synthetic.stress_flow(B, H, W, g)
    else:
        flow = synthetic.border_flow(B, H, W, g, margin=max(1, min(H, W) // 4), reach=40.0)
    ref = oracle.torch_warp(inp, flow)
    got = _warp(inp.to(_dev()), flow.to(_dev()), "cpu", "tma")
    assert_warp_close(got, ref, f"{shape} {kind} tma")
    assert torch.equal(got, _warp(inp.to(_dev()), flow.to(_dev()), "cpu", "gather"))


def test_tma_rejects_unaligned_rows():
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib
    x = torch.randn(1, 8, 16, 30, device=_dev())  # W % 4 != 0: TMA needs 16-byte rows
    fl = torch.zeros(1, 2, 16, 30, device=_dev())
    with pytest.raises(_lib.DeepSVCNativeError):
        d.warp_forward(x, fl, algo=_lib.WARP_TMA)
    d.warp_forward(x, fl, algo=_lib.WARP_AUTO)  # auto falls back to the gather kernel


def test_gather_and_auto_agree_bitwise():
    from deepsvc_b200 import synthetic
    g = torch.Generator().manual_seed(11)
    inp = torch.randn(1, 64, 128, 256, generator=g).to(_dev())
    flow = synthetic.smooth_flow(1, 128, 256, g).to(_dev())
    assert torch.equal(_warp(inp, flow, "cuda", "gather"), _warp(inp, flow, "cuda", "auto"))


def test_channels_last(oracle):
    from deepsvc_b200 import synthetic
    import deepsvc_b200 as d
    g = torch.Generator().manual_seed(5)
    inp = torch.randn(2, 64, 48, 80, generator=g)
    flow = synthetic.smooth_flow(2, 48, 80, g)
    ref = oracle.torch_warp(inp, flow)
    x = inp.to(_dev()).contiguous(memory_format=torch.channels_last)
    d.set_flow_arithmetic("cpu")
    try:
        out = d.torch_warp(x, flow.to(_dev()))
    finally:
        d.set_flow_arithmetic("cuda")
    assert out.is_contiguous(memory_format=torch.channels_last)
    assert_warp_close(out, ref, "nhwc")
    # strides the kernels cannot read in place: copied once (the reference's grid_sample accepts
    # them), an error in strict mode; a non-contiguous flow likewise
    from deepsvc_b200 import warp as Wm
    bad = torch.randn(1, 3, 8, 8).to(_dev()).contiguous(memory_format=torch.channels_last)
    fl = (torch.randn(1, 2, 4, 16, generator=g) * 2).to(_dev())
    got = d.torch_warp(bad[:, :, ::2], fl[:, :, :, ::2])
    assert_warp_close(got, oracle.torch_warp(bad[:, :, ::2].contiguous(), fl[:, :, :, ::2].contiguous()), "strided")
    Wm.STRICT_STRIDES = True
    try:
        with pytest.raises(RuntimeError, match="contiguous"):
            d.torch_warp(bad[:, :, ::2], torch.zeros(1, 2, 4, 8, device=_dev()))
    finally:
        Wm.STRICT_STRIDES = False
    # channels_last input through autograd: gradients come back in the input's memory format
    xg = x.clone().requires_grad_(True)
    fg = flow.to(_dev()).clone().requires_grad_(True)
    xr = inp.to(_dev()).clone().requires_grad_(True)
    fr = flow.to(_dev()).clone().requires_grad_(True)
    cot = torch.randn(2, 64, 48, 80, generator=g).to(_dev())
    d.torch_warp(xg, fg).backward(cot.contiguous(memory_format=torch.channels_last))
    d.torch_warp(xr, fr).backward(cot)
    assert xg.grad.is_contiguous(memory_format=torch.channels_last)
    assert torch.allclose(xg.grad, xr.grad, rtol=1e-4, atol=1e-5) and torch.allclose(fg.grad, fr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "warp_*.npz"))))
def test_backward_golden(path):
    """grad_input / grad_flow against autograd of the reference function (fixtures)."""
    import deepsvc_b200 as d
    g = np.load(path)
    inp = torch.from_numpy(g["input"]).to(_dev()).requires_grad_(True)
    flow = torch.from_numpy(g["flow"]).to(_dev()).requires_grad_(True)
    d.set_flow_arithmetic("cpu")
    try:
        out = d.torch_warp(inp, flow)
        gin, gflow = torch.autograd.grad(out, (inp, flow), torch.from_numpy(g["grad_out"]).to(_dev()))
    finally:
        d.set_flow_arithmetic("cuda")
    # gradients: 1e-4 relative to the tensor's scale (atomics reorder the fp32 sums)
    for got, ref, nm in ((gin, g["grad_input"], "grad_input"), (gflow, g["grad_flow"], "grad_flow")):
        ref = torch.from_numpy(ref)
        err = (got.cpu() - ref).abs().max().item()
        assert err <= 1e-4 * max(1.0, ref.abs().max().item()), f"{nm} {err}"


@pytest.mark.parametrize("need", [(True, True), (False, True), (True, False)])
def test_backward_vs_stock_torch_cuda(oracle, need):
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    g = torch.Generator().manual_seed(9)
    B, C, H, W = 2, 16, 40, 56
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow0 = synthetic.smooth_flow(B, H, W, g, sigma=6.0).to(_dev())
    # push some samples onto / past the border: gradient must vanish there
    flow0[:, 0, :, :3] = -10.0
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    res = []
    for fn in (oracle.torch_warp, d.torch_warp):
        inp = inp0.clone().requires_grad_(need[0])
        flow = flow0.clone().requires_grad_(need[1])
        out = fn(inp, flow)
        out.backward(gout)
        res.append((inp.grad, flow.grad))
    for a, b, nm in zip(res[1], res[0], ("grad_input", "grad_flow")):
        if b is None:
            assert a is None
            continue
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{nm} {err}"
    if need[1]:
        assert res[1][1][:, 0, :, :3].abs().max().item() == 0.0


def test_full_size_1080p_properties():
    """BASELINE config 2 size (1088x1920, 64 ch): size-independent checks -- an integer
    shift is reproduced exactly in the interior, and the op is idempotent under zero
    flow up to coordinate rounding."""
    dev = _dev()
    g = torch.Generator().manual_seed(2)
    inp = torch.randn(1, 64, 1088, 1920, generator=g).to(dev)
    flow = torch.zeros(1, 2, 1088, 1920, device=dev)
    flow[:, 0] = 5.0
    flow[:, 1] = -3.0
    import deepsvc_b200 as d
    out = d.torch_warp(inp, flow)
    ref = inp[:, :, :-3, 5:]
    got = out[:, :, 3:, :-5]
    assert (got - ref).abs().max().item() <= 5e-3  # coordinate rounding ~2e-4 px * gradient
    # checksum property: sum over taps of weights is 1 -> mean preserved for a shift
    assert abs(got.double().mean().item() - ref.double().mean().item()) < 1e-5


@pytest.mark.parametrize("cut", ["x", "y", "xy", "wild_corner"])
def test_tma_motion_boundary_quadrants(oracle, cut):
    """Tiles whose bounding box does not fit the staging box because a motion boundary
    crosses them: the staged kernel re-stages them as quadrants (each side of the boundary
    fits on its own), or gathers a quadrant that is itself wild.  Bit-identical to the
    gather kernel, within tolerance of the oracle."""
    from deepsvc_b200 import synthetic
    g = torch.Generator().manual_seed(11)
    B, C, H, Wd = 1, 20, 128, 256
    inp = torch.randn(B, C, H, Wd, generator=g)
    flow = torch.randn(B, 2, H, Wd, generator=g) * 0.3
    ys = torch.arange(H).view(1, H, 1).expand(B, H, Wd)
    xs = torch.arange(Wd).view(1, 1, Wd).expand(B, H, Wd)
    if "x" in cut:   # boundary in the middle of every 64-wide tile
        flow[:, 0] += torch.where((xs % 64) < 32, 30.0, -30.0)
    if "y" in cut:   # boundary in the middle of every 32-high tile
        flow[:, 1] += torch.where((ys % 32) < 16, -25.0, 25.0)
    if cut == "wild_corner":
        flow[:, :, :16, :32] = synthetic.stress_flow(B, 16, 32, g, sigma=40.0)
        flow[:, 0, 64:, 128:] += torch.where((xs[:, 64:, 128:] % 64) < 32, 30.0, -30.0)
    ref = oracle.torch_warp(inp, flow)
    got = _warp(inp.to(_dev()), flow.to(_dev()), "cpu", "tma")
    assert_warp_close(got, ref, f"motion boundary {cut}")
    assert torch.equal(got, _warp(inp.to(_dev()), flow.to(_dev()), "cpu", "gather"))


@pytest.mark.parametrize("kind", ["smooth", "stress"])
def test_tma_scheduler_state_left_zeroed(oracle, kind):
    """The staged kernel's scheduler state (workspace) is zero before and after every
    launch, also when every tile is unstageable (stress flow: every rectangle is gathered),
    and back-to-back launches on one workspace are bit-identical."""
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib, synthetic, warp as W
    g = torch.Generator().manual_seed(5)
    B, C, H, Wd = 2, 24, 96, 192
    inp = torch.randn(B, C, H, Wd, generator=g)
    flow = (synthetic.smooth_flow if kind == "smooth" else synthetic.stress_flow)(B, H, Wd, g)
    ref = oracle.torch_warp(inp, flow)
    x, f = inp.to(_dev()), flow.to(_dev())
    outs = [d.warp_forward(x, f, flow_mode=_lib.FLOW_TRUE_DIVIDE, algo=_lib.WARP_TMA) for _ in range(3)]
    torch.cuda.synchronize()
    ws = W.warp_workspace(x.device, B, H, Wd)
    assert int(ws.count_nonzero()) == 0
    assert_warp_close(outs[0], ref, f"scheduler {kind}")
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    assert torch.equal(outs[0], d.warp_forward(x, f, flow_mode=_lib.FLOW_TRUE_DIVIDE, algo=_lib.WARP_GATHER))


# ------------------------------------------------------------------ staged backward (warp_bwd_staged.cu)
def _bwd_algo(algo):
    from deepsvc_b200 import _lib
    _lib.check(_lib.load().dsvc_set_warp_bwd_algo(algo), "dsvc_set_warp_bwd_algo")


@pytest.mark.parametrize("need", [(True, True), (False, True), (True, False)])
@pytest.mark.parametrize("kind", ["smooth", "stress", "border"])
@pytest.mark.parametrize("shape", [(2, 16, 40, 64), (1, 64, 128, 192), (8, 64, 64, 64), (1, 9, 33, 100),
                                   (1, 8, 16, 68)])
@pytest.mark.parametrize("bwd", ["staged", "gather", "cell"])
def test_backward_staged_vs_stock_torch_cuda(oracle, shape, kind, need, bwd):
    """The staged / gather / cell-order backward kernels (forced) against autograd of the reference function on
    the GPU (its GPU branch, modules.py:44-62): gradients within 1e-4 of the tensor's scale."""
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib, synthetic
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H * 131 + W + C)
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow0 = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    res = []
    if bwd in ("gather", "cell") and not need[0]:
        pytest.skip("the gather / cell kernels produce grad_input; flow-only gradients use the other kernels")
    forced = {"staged": _lib.WARP_BWD_STAGED, "gather": _lib.WARP_BWD_GATHER, "cell": _lib.WARP_BWD_CELL}[bwd]
    for fn, algo in ((oracle.torch_warp, _lib.WARP_BWD_AUTO), (d.torch_warp, forced)):
        _bwd_algo(algo)
        try:
            inp = inp0.clone().requires_grad_(need[0])
            flow = flow0.clone().requires_grad_(need[1])
            fn(inp, flow).backward(gout)
            res.append((inp.grad, flow.grad))
        finally:
            _bwd_algo(_lib.WARP_BWD_AUTO)
    for a, b, nm in zip(res[1], res[0], ("grad_input", "grad_flow")):
        if b is None:
            assert a is None
            continue
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{nm} {err}"


def test_backward_staged_matches_direct_kernel():
    """Staged and per-pixel kernels on the cfg3 shape (B=8, 64 ch, 256x256): same gradients up
    to fp32 summation order; grad_input of a constant grad_out sums to the number of taps."""
    from deepsvc_b200 import _lib, synthetic
    from deepsvc_b200.warp import warp_backward
    B, C, H, W = 8, 64, 256, 256
    g = torch.Generator().manual_seed(5)
    inp = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow = synthetic.smooth_flow(B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    out = {}
    for name, algo in (("direct", _lib.WARP_BWD_DIRECT), ("staged", _lib.WARP_BWD_STAGED),
                       ("gather", _lib.WARP_BWD_GATHER), ("cell", _lib.WARP_BWD_CELL)):
        _bwd_algo(algo)
        try:
            out[name] = warp_backward(gout, inp, flow, True, True)
            ones = warp_backward(torch.ones_like(gout), inp, flow, True, False)[0]
        finally:
            _bwd_algo(_lib.WARP_BWD_AUTO)
        # bilinear weights of a pixel sum to 1: every plane of grad_input sums to H*W
        s = ones.double().sum((2, 3))
        assert (s - H * W).abs().max().item() <= 1e-3 * H * W, name
    for which in ("staged", "gather", "cell"):
        for a, b, nm in zip(out[which], out["direct"], ("grad_input", "grad_flow")):
            err = (a - b).abs().max().item()
            assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{which} {nm} {err}"


@pytest.mark.parametrize("kind,want", [("smooth", 1), ("gentle", 1), ("stress", 2)])
def test_backward_scout_picks_kernel_per_launch(oracle, kind, want, monkeypatch):
    """With a workspace too small for the cell tables the default backward samples the flow on the
    device and runs the staged kernel when at least half the 64x16 tiles fit its 96x32 staging box,
    the per-pixel kernel otherwise; the decision word (workspace[0]) is 1 / 2 and the gradients
    match autograd of the reference."""
    from deepsvc_b200 import synthetic, warp as w
    monkeypatch.setattr(w, "CELL_MIN_CHANNELS", 1 << 30)   # small workspace: no cell tables
    w._bwd_ws_cache.clear()
    B, C, H, W = 1, 64, 272, 480
    g = torch.Generator().manual_seed(77)
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow0 = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    try:
        gin, gflow = w.warp_backward(gout, inp0, flow0, True, True)
        ws = w._bwd_workspace(_dev(), B, H, W)
        state = ws[:16].view(torch.int32).tolist()
    finally:
        w._bwd_ws_cache.clear()
    assert state[0] == want and state[1] == 0 and state[2] == 0, state
    inp, flow = inp0.clone().requires_grad_(True), flow0.clone().requires_grad_(True)
    oracle.torch_warp(inp, flow).backward(gout)
    for a, b in ((gin, inp.grad), (gflow, flow.grad)):
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("kind", ["smooth", "stress", "border", "gentle"])
@pytest.mark.parametrize("shape", [(1, 64, 272, 480), (2, 8, 100, 132), (1, 16, 31, 31), (1, 8, 7, 200), (3, 9, 64, 33)])
def test_backward_cell_order_is_the_default_for_wide_warps(oracle, shape, kind):
    """C >= 8 with the full workspace: the cell-order kernel (csrc/warp_bwd_cell.cu) runs by default --
    grad_input is NOT pre-zeroed by anyone (poisoned here) and every element must be written; both
    gradients against autograd of the reference on the GPU, for flows that fill cells unevenly
    (stress: Poisson occupancy; border: hundreds of pixels clamped into one cell -> overflow list)."""
    from deepsvc_b200 import _lib, synthetic, warp as w
    B, C, H, W = shape
    g = torch.Generator().manual_seed(sum(shape))
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow0 = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    lib = _lib.load()
    lin_x, lin_y = w._base_grids(_dev(), H, W)
    sx, sy, inv_sx, inv_sy = w._scales(H, W)
    ws = w._bwd_workspace(_dev(), B, H, W, big=True)
    assert ws.numel() >= lib.dsvc_warp_bwd_cell_workspace_bytes(B, H, W)
    gin = torch.full_like(inp0, float("nan"))
    gflow = torch.full_like(flow0, float("nan"))
    _lib.check(lib.dsvc_warp_bwd_ws_f32(gout.data_ptr(), inp0.data_ptr(), flow0.data_ptr(), gin.data_ptr(), gflow.data_ptr(),
                                        B, C, H, W, lin_x.data_ptr(), lin_y.data_ptr(), sx, sy, inv_sx, inv_sy,
                                        _lib.FLOW_MUL_RECIPROCAL, _lib.LAYOUT_NCHW, ws.data_ptr(), ws.numel(),
                                        torch.cuda.current_stream(_dev()).cuda_stream), "dsvc_warp_bwd_ws_f32")
    inp, flow = inp0.clone().requires_grad_(True), flow0.clone().requires_grad_(True)
    oracle.torch_warp(inp, flow).backward(gout)
    assert torch.isfinite(gin).all() and torch.isfinite(gflow).all()
    for a, b, nm in ((gin, inp.grad, "grad_input"), (gflow, flow.grad, "grad_flow")):
        err = (a - b).abs().max().item()
        assert err <= 1e-4 * max(1.0, b.abs().max().item()), f"{nm} {err}"


def test_backward_staged_collapsed_flow():
    """Flow that collapses a whole tile onto one source column (a destination element with 64
    taps per row: its run spans several threads' shares and is summed piecewise through
    shared-memory atomics) must still be right."""
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib
    B, C, H, W = 1, 8, 32, 128
    g = torch.Generator().manual_seed(2)
    inp0 = torch.randn(B, C, H, W, generator=g).to(_dev())
    flow0 = torch.zeros(B, 2, H, W)
    flow0[:, 0] = 40.3 - torch.arange(W, dtype=torch.float32)[None, None, :]  # every x -> 40.3
    flow0 = flow0.to(_dev())
    gout = torch.randn(B, C, H, W, generator=g).to(_dev())
    res = []
    for algo in (_lib.WARP_BWD_DIRECT, _lib.WARP_BWD_STAGED, _lib.WARP_BWD_GATHER, _lib.WARP_BWD_CELL):
        _bwd_algo(algo)
        try:
            inp = inp0.clone().requires_grad_(True)
            flow = flow0.clone().requires_grad_(True)
            d.torch_warp(inp, flow).backward(gout)
            res.append((inp.grad, flow.grad))
        finally:
            _bwd_algo(_lib.WARP_BWD_AUTO)
    for r in res[1:]:
        for a, b in zip(r, res[0]):
            assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())


def test_full_size_4k_forward_and_backward_properties(oracle):
    """BASELINE config 5 size (2176x3840, 64 ch = 2.14 GB per tensor): forward against the stock
    torch CUDA branch of the reference function on a smooth flow, and two size-independent
    backward properties of the staged kernel (plane sums of grad_input for a constant grad_out;
    grad_flow against stock autograd on a 16-channel slice)."""
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from deepsvc_b200.warp import warp_backward
    dev = _dev()
    H, W = 2176, 3840
    g = torch.Generator().manual_seed(4)
    flow = synthetic.smooth_flow(1, H, W, g).to(dev)
    inp = torch.randn(1, 64, H, W, generator=g).to(dev)
    out = d.torch_warp(inp, flow)
    ref = oracle.torch_warp(inp, flow)
    assert (out - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    del ref, out
    torch.cuda.empty_cache()
    gin, _ = warp_backward(torch.ones_like(inp), inp, flow, True, False)
    s = gin.double().sum((2, 3))
    assert (s - H * W).abs().max().item() <= 1e-3 * H * W
    del gin
    torch.cuda.empty_cache()
    x16 = inp[:, :16].contiguous().requires_grad_(True)
    f = flow.clone().requires_grad_(True)
    go = torch.randn(1, 16, H, W, generator=g).to(dev)
    oracle.torch_warp(x16, f).backward(go)
    gi_ref, gf_ref = x16.grad, f.grad
    gi, gf = warp_backward(go, x16.detach(), flow, True, True)
    assert (gi - gi_ref).abs().max().item() <= 1e-4 * max(1.0, gi_ref.abs().max().item())
    assert (gf - gf_ref).abs().max().item() <= 1e-4 * max(1.0, gf_ref.abs().max().item())


@pytest.mark.parametrize("kind", ["smooth", "stress", "border"])
@pytest.mark.parametrize("shape", [(1, 64, 3, 128, 192), (2, 16, 3, 100, 132), (1, 8, 5, 64, 64), (1, 9, 2, 33, 100),
                                   (1, 3, 3, 40, 64), (1, 64, 3, 1088, 1920)])
def test_two_tensors_one_flow_bit_identical(shape, kind):
    """dsvc_warp_fwd2_f32 (frame planes riding on the feature warp's staged launch, or two plain
    launches for shapes the staged kernel does not take) == two separate torch_warp calls, bit for bit;
    the scheduler state is left zeroed."""
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from deepsvc_b200.warp import warp_forward, warp_forward2, warp_workspace
    B, Ca, Cb, H, W = shape
    g = torch.Generator().manual_seed(Ca * 7 + H + W)
    a = torch.randn(B, Ca, H, W, generator=g).to(_dev())
    b = torch.rand(B, Cb, H, W, generator=g).to(_dev())
    flow = synthetic.make_flow(kind, B, H, W, g).to(_dev())
    for _ in range(2):
        oa, ob = warp_forward2(a, b, flow)
        assert torch.equal(oa, warp_forward(a, flow))
        assert torch.equal(ob, warp_forward(b, flow))
    assert int(warp_workspace(_dev(), B, H, W).abs().sum().item()) == 0
    with pytest.raises(RuntimeError):
        warp_forward2(a.requires_grad_(True), b, flow)


def test_few_channel_kernel_claims_tiles_and_leaves_scheduler_zeroed(oracle):
    """The 3-ch kernel claims its tiles from the workspace's scheduler words (so that it can share
    the SMs with the persistent feature warp): same bits as the static grid-stride order (no
    workspace), and the words are zero again after every launch."""
    from deepsvc_b200 import _lib, synthetic
    from deepsvc_b200.warp import _base_grids, _scales
    lib = _lib.load()
    g = torch.Generator().manual_seed(12)
    for (B, C, H, W) in ((1, 3, 272, 480), (2, 3, 70, 100), (1, 1, 9, 40)):
        inp = torch.randn(B, C, H, W, generator=g).to(_dev())
        flow = synthetic.smooth_flow(B, H, W, g).to(_dev())
        lx, ly = _base_grids(_dev(), H, W)
        sx, sy, isx, isy = _scales(H, W)
        ws = torch.zeros(64, dtype=torch.uint8, device=_dev())
        outs = []
        for wsp, wsn in ((ws.data_ptr(), ws.numel()), (None, 0), (ws.data_ptr(), ws.numel())):
            out = torch.empty_like(inp)
            _lib.check(lib.dsvc_warp_fwd_f32(inp.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, H, W, lx.data_ptr(),
                                             ly.data_ptr(), sx, sy, isx, isy, _lib.FLOW_MUL_RECIPROCAL, _lib.LAYOUT_NCHW,
                                             _lib.WARP_AUTO, wsp, wsn, torch.cuda.current_stream().cuda_stream), "warp")
            torch.cuda.synchronize()
            assert int(ws.view(torch.int32).abs().sum()) == 0
            outs.append(out)
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
        assert_warp_close(outs[0], oracle.torch_warp(inp, flow), "few-channel, claimed tiles")
