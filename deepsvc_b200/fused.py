"""Fusions on either side of the few-channel warps.

``spynet_level_warp`` replaces the three statements of one SpyNet level
(``modules.py:163-168``)::

    flow_up = bilinearupsacling(flow) * 2.0
    warped  = torch_warp(im2, flow_up)

and ``warp_with_mse`` the motion-compensation pair of ``video_model.py:37-38``::

    warped_frame = torch_warp(ref_frame, recon_mv)
    warp_loss    = torch.mean((warped_frame - curr_frame).pow(2))

Each is one launch of ``dsvc_warp_fused_f32`` (``csrc/warp_fused.cu``).  All fusions are
differentiable: the forward is the fused launch, the backward runs the warp backward kernel
(``dsvc_warp_bwd_ws_f32``) plus the few elementwise / interpolation gradients around it, with
the gradients of the unfused chain (``tests/test_gpu_fused.py``).  No CPU path.
"""
import torch

from . import _lib
from . import warp as _warp
from .entropy import bits_finalize


_finalize_consts = {}


def _check(name, *ts):
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"deepsvc_b200.{name}: contiguous fp32 CUDA tensors required (no CPU fallback)")


def _wants_grad(*ts):
    return torch.is_grad_enabled() and any(t.requires_grad for t in ts)


def _launch(inp, flow, flow_coarse, flow_up, target, partials, out):
    B, C, H, W = inp.shape
    if C > 4:
        raise RuntimeError("deepsvc_b200 fused warps: C <= 4 (frames, SpyNet pyramid levels)")
    lin_x, lin_y = _warp._base_grids(inp.device, H, W)
    sx, sy, inv_sx, inv_sy = _warp._scales(H, W)
    with torch.cuda.device(inp.device):
        err = _lib.load().dsvc_warp_fused_f32(
            inp.data_ptr(), _lib.ptr(flow), _lib.ptr(flow_coarse), _lib.ptr(flow_up), _lib.ptr(target),
            _lib.ptr(partials), out.data_ptr(), B, C, H, W, lin_x.data_ptr(), lin_y.data_ptr(),
            sx, sy, inv_sx, inv_sy, _warp._flow_mode, _lib.stream_ptr(inp.device))
    _lib.check(err, "dsvc_warp_fused_f32")


def _spynet_level_launch(im2, flow):
    B, C, H, W = im2.shape
    flow_up = torch.empty(B, 2, H, W, dtype=torch.float32, device=im2.device)
    out = torch.empty_like(im2)
    if im2.numel():
        _launch(im2, None, flow, flow_up, None, None, out)
    return flow_up, out


class _SpynetLevelFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im2, flow):
        flow_up, out = _spynet_level_launch(im2, flow)
        ctx.save_for_backward(im2, flow_up)
        ctx.flow_mode = _warp._flow_mode
        return flow_up, out

    @staticmethod
    def backward(ctx, g_flow_up, g_out):
        im2, flow_up = ctx.saved_tensors
        g_im2 = g_fu = None
        if g_out is not None:
            g_im2, g_fu = _warp.warp_backward(g_out, im2, flow_up, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                              ctx.flow_mode)
        g_flow = None
        if ctx.needs_input_grad[1]:
            total = g_fu if g_flow_up is None else (g_flow_up if g_fu is None else g_flow_up + g_fu)
            if total is not None:
                # adjoint of `F.interpolate(flow, x2, bilinear, align_corners=False) * 2.0` (modules.py:107-112)
                B, _, H, W = flow_up.shape
                g_flow = torch.ops.aten.upsample_bilinear2d_backward(total.contiguous() * 2.0, [H, W],
                                                                     [B, 2, H // 2, W // 2], False, None, None)
        return g_im2, g_flow


def spynet_level_warp(im2: torch.Tensor, flow: torch.Tensor):
    """(flow_up [B,2,H,W], warped [B,C,H,W]) from im2 [B,C,H,W] and the previous level's
    flow [B,2,H/2,W/2] (``modules.py:163-168``).  Differentiable in both arguments."""
    _check("spynet_level_warp", im2, flow)
    B, C, H, W = im2.shape
    if flow.shape != (B, 2, H // 2, W // 2) or H % 2 or W % 2:
        raise RuntimeError("deepsvc_b200.spynet_level_warp: flow must be [B,2,H/2,W/2] of an even-sized image")
    if _wants_grad(im2, flow):
        return _SpynetLevelFn.apply(im2, flow)
    return _spynet_level_launch(im2, flow)


class _WarpMseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref_frame, flow, curr_frame):
        out, loss = _warp_with_mse_launch(ref_frame, flow, curr_frame)
        ctx.save_for_backward(ref_frame, flow, curr_frame, out)
        ctx.flow_mode = _warp._flow_mode
        return out, loss

    @staticmethod
    def backward(ctx, g_out, g_loss):
        ref_frame, flow, curr_frame, out = ctx.saved_tensors
        # d loss / d warped = 2 (warped - cur) / N  (video_model.py:38)
        g = None
        g_cur = None
        if g_loss is not None:
            d = (out - curr_frame) * (2.0 / out.numel()) * g_loss.to(torch.float32)
            g = d
            if ctx.needs_input_grad[2]:
                g_cur = -d
        if g_out is not None:
            g = g_out if g is None else g + g_out
        g_ref = g_flow = None
        if g is not None and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            g_ref, g_flow = _warp.warp_backward(g, ref_frame, flow, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                                ctx.flow_mode)
        return g_ref, g_flow, g_cur


def warp_with_mse(ref_frame: torch.Tensor, flow: torch.Tensor, curr_frame: torch.Tensor):
    """(warped_frame, warp_loss) of ``video_model.py:37-38``; warp_loss is a 0-d fp64 device
    tensor (fixed-order sum, bit-identical reruns).  Differentiable in all three arguments."""
    _check("warp_with_mse", ref_frame, flow, curr_frame)
    B, C, H, W = ref_frame.shape
    if flow.shape != (B, 2, H, W) or curr_frame.shape != ref_frame.shape:
        raise RuntimeError("deepsvc_b200.warp_with_mse: shape mismatch")
    if _wants_grad(ref_frame, flow, curr_frame):
        return _WarpMseFn.apply(ref_frame, flow, curr_frame)
    return _warp_with_mse_launch(ref_frame, flow, curr_frame)


def _warp_with_mse_launch(ref_frame, flow, curr_frame):
    B, C, H, W = ref_frame.shape
    dev = ref_frame.device
    n = _lib.load().dsvc_warp_fused_slots(B, H, W)
    partials = torch.empty(n, dtype=torch.float64, device=dev)
    out = torch.empty_like(ref_frame)
    _launch(ref_frame, flow, None, None, curr_frame, partials, out)
    key = (dev, n, ref_frame.numel())
    consts = _finalize_consts.get(key)
    if consts is None:  # built once per shape: no host-to-device copy on the hot path
        consts = (torch.tensor([0, n], dtype=torch.int32, device=dev),
                  torch.full((1,), 1.0 / ref_frame.numel(), dtype=torch.float64, device=dev))
        _finalize_consts[key] = consts
    return out, bits_finalize(partials, consts[0], consts[1])[0]


class _BlendFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, warped, pred):
        ctx.save_for_backward(w, warped, pred)
        return _blend_launch(w, warped, pred)

    @staticmethod
    def backward(ctx, g):
        w, warped, pred = ctx.saved_tensors
        return (g * (warped - pred) if ctx.needs_input_grad[0] else None,
                g * w if ctx.needs_input_grad[1] else None,
                g * (1 - w) if ctx.needs_input_grad[2] else None)


def mc_blend(w: torch.Tensor, warped: torch.Tensor, pred: torch.Tensor) -> torch.Tensor:
    """``w * warped + (1 - w) * pred`` (``modules.py:436``) in one pass, bit-identical to the
    reference's expression.  Differentiable in all three arguments."""
    _check("mc_blend", w, warped, pred)
    if not (w.shape == warped.shape == pred.shape):
        raise RuntimeError("deepsvc_b200.mc_blend: shape mismatch")
    if _wants_grad(w, warped, pred):
        return _BlendFn.apply(w, warped, pred)
    return _blend_launch(w, warped, pred)


def _blend_launch(w, warped, pred):
    out = torch.empty_like(w)
    with torch.cuda.device(w.device):
        err = _lib.load().dsvc_blend_f32(w.data_ptr(), warped.data_ptr(), pred.data_ptr(), out.data_ptr(),
                                         w.numel(), _lib.stream_ptr(w.device))
    _lib.check(err, "dsvc_blend_f32")
    return out


class _LrpAddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_hat, lrp):
        ctx.save_for_backward(lrp)
        return _lrp_add_launch(y_hat, lrp, None)

    @staticmethod
    def backward(ctx, g):
        (lrp,) = ctx.saved_tensors
        g_lrp = None
        if ctx.needs_input_grad[1]:
            g = g.contiguous()
            g_lrp = torch.empty_like(lrp)
            with torch.cuda.device(lrp.device):
                err = _lib.load().dsvc_lrp_add_bwd_f32(g.data_ptr(), lrp.data_ptr(), g_lrp.data_ptr(), lrp.numel(),
                                                       _lib.stream_ptr(lrp.device))
            _lib.check(err, "dsvc_lrp_add_bwd_f32")
        return (g if ctx.needs_input_grad[0] else None), g_lrp


def _lrp_add_launch(y_hat, lrp, out):
    out = torch.empty_like(y_hat) if out is None else out
    with torch.cuda.device(y_hat.device):
        err = _lib.load().dsvc_lrp_add_f32(y_hat.data_ptr(), lrp.data_ptr(), out.data_ptr(), y_hat.numel(),
                                           _lib.stream_ptr(y_hat.device))
    _lib.check(err, "dsvc_lrp_add_f32")
    return out


def lrp_add(y_hat: torch.Tensor, lrp: torch.Tensor, inplace: bool = False) -> torch.Tensor:
    """``y_hat + 0.5 * tanh(lrp)`` -- ``image_model.py:185-188`` (``lrp = 0.5 * torch.tanh(lrp);
    y_hat_slice += lrp``) in one launch instead of three, bit-identical; differentiable in both
    arguments.  ``inplace=True`` writes into ``y_hat`` like the reference's ``+=`` (no autograd)."""
    for t in (y_hat, lrp):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError("deepsvc_b200.lrp_add: fp32 CUDA tensors required (no CPU fallback)")
    if y_hat.shape != lrp.shape:
        raise RuntimeError("deepsvc_b200.lrp_add: shape mismatch")
    if not (y_hat.is_contiguous() and lrp.is_contiguous()):
        y_hat_c, lrp = y_hat.contiguous(), lrp.contiguous()
        if inplace:
            return y_hat.copy_(_lrp_add_launch(y_hat_c, lrp, None))
        y_hat = y_hat_c
    if inplace:
        return _lrp_add_launch(y_hat, lrp, y_hat)
    if torch.is_grad_enabled() and (y_hat.requires_grad or lrp.requires_grad):
        return _LrpAddFn.apply(y_hat, lrp)
    return _lrp_add_launch(y_hat, lrp, None)


__all__ = ["spynet_level_warp", "warp_with_mse", "mc_blend", "lrp_add"]
