"""Drop-in replacement for the reference's ``torch_warp`` (``modules.py:25-62``).

Same call signature and tensor layout: ``torch_warp(tensorInput[B,C,H,W],
tensorFlow[B,2,H,W]) -> [B,C,H,W]``, fp32, differentiable in both arguments.  The
work is one hand-written sm_100a launch (``csrc/warp_*.cu``) instead of the
reference's 5 preparation launches + ATen ``grid_sampler_2d``; the base-grid caches
of ``modules.py:21-22`` are replaced by two tiny ``linspace`` tables per (device, H, W).

There is no CPU path: a non-CUDA or non-fp32 tensor raises.
"""
import numpy as np
import torch

from . import _lib

# names kept for import compatibility with ``modules.py:21-22`` (unused)
Backward_tensorGrid = [{} for _ in range(8)]
Backward_tensorGrid_cpu = {}

_flow_mode = _lib.FLOW_MUL_RECIPROCAL
_algo = _lib.WARP_AUTO
_lin_cache = {}


def set_flow_arithmetic(which: str) -> None:
    """Select which branch of the reference the flow scaling reproduces bit for bit:
    "cuda" (``modules.py:54-55`` executed by ATen as a multiply by the fp32 reciprocal;
    the default, since this op replaces the reference's CUDA branch) or "cpu"
    (``modules.py:36-37``, a true division)."""
    global _flow_mode
    if which not in ("cuda", "cpu"):
        raise ValueError('flow arithmetic must be "cuda" or "cpu"')
    _flow_mode = _lib.FLOW_MUL_RECIPROCAL if which == "cuda" else _lib.FLOW_TRUE_DIVIDE


def set_warp_algorithm(which: str) -> None:
    """"auto" (default), "gather" or "tma"."""
    global _algo
    _algo = {"auto": _lib.WARP_AUTO, "gather": _lib.WARP_GATHER, "tma": _lib.WARP_TMA}[which]


def _base_grids(device, H, W):
    key = (device, H, W)
    g = _lin_cache.get(key)
    if g is None:
        # modules.py:47-50: the base grid is always computed by CPU linspace, then copied
        g = (torch.linspace(-1.0, 1.0, W).to(device), torch.linspace(-1.0, 1.0, H).to(device))
        _lin_cache[key] = g
    return g


_ws_cache = {}


def warp_workspace(device, B, H, W, private=False):
    """Work list for the staged kernel (``dsvc_warp_fwd_f32``'s `workspace`): zero-filled
    once, left zero-filled by every launch.  One buffer per (device, stream) is cached and
    reused by eager calls (launches on one stream are ordered); ``private=True`` returns
    a fresh buffer for a caller that binds it into its own launch sequence / CUDA graph."""
    n = _lib.load().dsvc_warp_workspace_bytes(B, H, W)
    if private or torch.cuda.is_current_stream_capturing():
        # under a user's graph capture the zero-fill is captured with the launch
        return torch.zeros(n, dtype=torch.uint8, device=device)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.zeros(max(n, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def _scales(H, W):
    sx = np.float32((W - 1.0) / 2.0)
    sy = np.float32((H - 1.0) / 2.0)
    with np.errstate(divide="ignore"):
        inv_sx = np.float32(1.0) / sx  # ATen div_true_kernel_cuda: a * (1 / b), opmath fp32
        inv_sy = np.float32(1.0) / sy
    return float(sx), float(sy), float(inv_sx), float(inv_sy)


def _check(inp, flow):
    if not (inp.is_cuda and flow.is_cuda):
        raise RuntimeError("deepsvc_b200.torch_warp: CUDA tensors required (no CPU fallback)")
    if inp.device != flow.device:
        raise RuntimeError("deepsvc_b200.torch_warp: input and flow are on different devices")
    if inp.dtype != torch.float32 or flow.dtype != torch.float32:
        raise RuntimeError("deepsvc_b200.torch_warp: fp32 tensors required")
    if inp.dim() != 4 or flow.dim() != 4 or flow.size(1) != 2:
        raise RuntimeError("deepsvc_b200.torch_warp: expected input [B,C,H,W] and flow [B,2,H,W]")
    if inp.size(0) != flow.size(0) or inp.shape[2:] != flow.shape[2:]:
        raise RuntimeError(
            f"deepsvc_b200.torch_warp: shape mismatch {tuple(inp.shape)} vs {tuple(flow.shape)}")


def _layout_of(inp):
    if inp.is_contiguous():
        return _lib.LAYOUT_NCHW
    if inp.is_contiguous(memory_format=torch.channels_last) and inp.size(1) % 4 == 0:
        return _lib.LAYOUT_NHWC
    raise RuntimeError("deepsvc_b200.torch_warp: input must be contiguous NCHW, or "
                       "channels_last with C % 4 == 0 (no silent layout copies)")


STRICT_STRIDES = False  # True: raise instead of copying a layout the kernels cannot read in place


def _dense(inp, flow):
    """The reference accepts any strides (``F.grid_sample`` does); the kernels read dense NCHW
    (or channels_last with C % 4 == 0).  Anything else is copied once, as the stock op's own
    ``.contiguous()`` would, unless STRICT_STRIDES asks for an error."""
    if not flow.is_contiguous():
        if STRICT_STRIDES:
            raise RuntimeError("deepsvc_b200.torch_warp: flow must be contiguous NCHW")
        flow = flow.contiguous()
    if not (inp.is_contiguous() or (inp.is_contiguous(memory_format=torch.channels_last)
                                    and inp.size(1) % 4 == 0)):
        if STRICT_STRIDES:
            raise RuntimeError("deepsvc_b200.torch_warp: input must be contiguous NCHW, or "
                               "channels_last with C % 4 == 0")
        inp = inp.contiguous()
    return inp, flow


def _drop_workspace(device):
    """A failed launch may leave the scheduler words non-zero: forget the cached buffers."""
    for key in [k for k in _ws_cache if k[0] == device]:
        del _ws_cache[key]


def warp_forward(inp, flow, flow_mode=None, algo=None):
    """Raw forward launch (no autograd)."""
    _check(inp, flow)
    inp, flow = _dense(inp, flow)
    layout = _layout_of(inp)
    B, C, H, W = inp.shape
    out = torch.empty_like(inp)  # preserves NCHW / channels_last
    if out.numel() == 0:
        return out
    lin_x, lin_y = _base_grids(inp.device, H, W)
    sx, sy, inv_sx, inv_sy = _scales(H, W)
    lib = _lib.load()
    ws = warp_workspace(inp.device, B, H, W) if layout == _lib.LAYOUT_NCHW else None
    with torch.cuda.device(inp.device):
        err = lib.dsvc_warp_fwd_f32(
            inp.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, H, W,
            lin_x.data_ptr(), lin_y.data_ptr(), sx, sy, inv_sx, inv_sy,
            _flow_mode if flow_mode is None else flow_mode, layout,
            _algo if algo is None else algo, _lib.ptr(ws), 0 if ws is None else ws.numel(),
            _lib.stream_ptr(inp.device))
    if err:
        _drop_workspace(inp.device)
    _lib.check(err, "dsvc_warp_fwd_f32")
    return out


def warp_forward2(inp_a, inp_b, flow, flow_mode=None):
    """(torch_warp(inp_a, flow), torch_warp(inp_b, flow)) in one launch (no autograd): two tensors
    that share a flow -- ``video_model.py:37`` and ``modules.py:429`` both warp by ``recon_mv``.
    Bit-identical to two ``warp_forward`` calls."""
    _check(inp_a, flow)
    _check(inp_b, flow)
    if not (inp_a.is_contiguous() and inp_b.is_contiguous()):
        raise RuntimeError("deepsvc_b200.warp_forward2: contiguous NCHW inputs required")
    if (inp_a.requires_grad or inp_b.requires_grad or flow.requires_grad) and torch.is_grad_enabled():
        raise RuntimeError("deepsvc_b200.warp_forward2: inference-only (use torch_warp for training)")
    B, Ca, H, W = inp_a.shape
    Cb = inp_b.shape[1]
    out_a, out_b = torch.empty_like(inp_a), torch.empty_like(inp_b)
    if out_a.numel() == 0 or out_b.numel() == 0:
        return warp_forward(inp_a, flow, flow_mode), warp_forward(inp_b, flow, flow_mode)
    lin_x, lin_y = _base_grids(inp_a.device, H, W)
    sx, sy, inv_sx, inv_sy = _scales(H, W)
    ws = warp_workspace(inp_a.device, B, H, W)
    with torch.cuda.device(inp_a.device):
        err = _lib.load().dsvc_warp_fwd2_f32(
            inp_a.data_ptr(), inp_b.data_ptr(), flow.data_ptr(), out_a.data_ptr(), out_b.data_ptr(),
            B, Ca, Cb, H, W, lin_x.data_ptr(), lin_y.data_ptr(), sx, sy, inv_sx, inv_sy,
            _flow_mode if flow_mode is None else flow_mode, _lib.ptr(ws), ws.numel(),
            _lib.stream_ptr(inp_a.device))
    if err:
        _drop_workspace(inp_a.device)
    _lib.check(err, "dsvc_warp_fwd2_f32")
    return out_a, out_b


def warp_backward(grad_out, inp, flow, need_input_grad=True, need_flow_grad=True, flow_mode=None):
    """Raw backward launch: returns (grad_input | None, grad_flow | None)."""
    _check(inp, flow)
    if not flow.is_contiguous():
        flow = flow.contiguous()
    if not inp.is_contiguous():
        # channels_last forward inputs: the backward kernels are NCHW (one copy; the gradient
        # comes back in the input's memory format)
        gin, gflow = warp_backward(grad_out, inp.contiguous(), flow, need_input_grad, need_flow_grad, flow_mode)
        if gin is not None:
            gin = gin.contiguous(memory_format=torch.channels_last) if inp.is_contiguous(
                memory_format=torch.channels_last) else gin
        return gin, gflow
    grad_out = grad_out.contiguous()
    B, C, H, W = inp.shape
    # neither gradient needs initialising: the gather kernel writes every element of grad_input
    # once, the other kernels zero-fill it inside dsvc_warp_bwd_ws_f32
    gin = torch.empty_like(inp) if need_input_grad else None
    gflow = torch.empty_like(flow) if need_flow_grad else None
    if inp.numel() == 0 or not (need_input_grad or need_flow_grad):
        return gin, gflow
    lin_x, lin_y = _base_grids(inp.device, H, W)
    sx, sy, inv_sx, inv_sy = _scales(H, W)
    lib = _lib.load()
    ws = _bwd_workspace(inp.device, B, H, W, big=C >= CELL_MIN_CHANNELS) if need_input_grad else None
    with torch.cuda.device(inp.device):
        err = lib.dsvc_warp_bwd_ws_f32(
            grad_out.data_ptr(), inp.data_ptr(), flow.data_ptr(), _lib.ptr(gin), _lib.ptr(gflow),
            B, C, H, W, lin_x.data_ptr(), lin_y.data_ptr(), sx, sy, inv_sx, inv_sy,
            _flow_mode if flow_mode is None else flow_mode, _lib.LAYOUT_NCHW,
            _lib.ptr(ws), 0 if ws is None else ws.numel(), _lib.stream_ptr(inp.device))
    _lib.check(err, "dsvc_warp_bwd_ws_f32")
    return gin, gflow


_bwd_ws_cache = {}
CELL_MIN_CHANNELS = 8   # the cell-order backward (csrc/warp_bwd_cell.cu) amortises its tables over the channels


def _bwd_workspace(device, B, H, W, big=False):
    """Per-tile flag bytes of the gather backward (written before they are read inside one
    call): one buffer per (device, stream), a fresh one under CUDA-graph capture."""
    n = _lib.load().dsvc_warp_bwd_workspace_bytes(B, H, W)
    if big:
        n = max(n, _lib.load().dsvc_warp_bwd_cell_workspace_bytes(B, H, W))
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(n, dtype=torch.uint8, device=device)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _bwd_ws_cache.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 14), dtype=torch.uint8, device=device)
        _bwd_ws_cache[key] = ws
    return ws


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, flow):
        ctx.save_for_backward(inp, flow)
        ctx.flow_mode = _flow_mode
        return warp_forward(inp, flow)

    @staticmethod
    def backward(ctx, grad_out):
        inp, flow = ctx.saved_tensors
        gin, gflow = warp_backward(grad_out, inp, flow, ctx.needs_input_grad[0],
                                   ctx.needs_input_grad[1], ctx.flow_mode)
        return gin, gflow


def torch_warp(tensorInput: torch.Tensor, tensorFlow: torch.Tensor) -> torch.Tensor:
    """Backward bilinear warp, border clamp, align_corners=True (``modules.py:25-62``)."""
    ops = _lib.torch_ops()
    if ops is not None and _algo == _lib.WARP_AUTO and not STRICT_STRIDES:
        # one dispatcher hop: checks, output allocation and autograd live in csrc_torch/ops.cpp
        return ops.torch_warp(tensorInput, tensorFlow, _flow_mode)
    if torch.is_grad_enabled() and (tensorInput.requires_grad or tensorFlow.requires_grad):
        _check(tensorInput, tensorFlow)
        return _WarpFn.apply(tensorInput, tensorFlow)
    return warp_forward(tensorInput, tensorFlow)


__all__ = ["torch_warp", "warp_forward", "warp_backward", "set_flow_arithmetic",
           "set_warp_algorithm", "Backward_tensorGrid", "Backward_tensorGrid_cpu"]
