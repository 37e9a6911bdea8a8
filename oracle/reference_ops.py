"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by ``deepsvc_b200``.

CPU (or, on the GPU box, stock-torch CUDA) restatement of the DeepSVC P-frame
warp + entropy hot path, used as the parity checker in ``tests/``, by
``__graft_entry__.smoke()`` and as ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` arm.

* ``torch_warp``      restates ``/root/reference/modules.py:25-62`` (both device
                      branches: base grid from CPU ``torch.linspace``, flow divided
                      by ``(W-1)/2`` / ``(H-1)/2``, ``F.grid_sample`` bilinear /
                      border / align_corners=True).  Pinned: validated bit-for-bit
                      against the reference's own function imported from
                      ``/root/reference`` (``oracle/make_golden.py``,
                      ``tests/test_oracle_cpu.py``) and frozen in ``tests/golden``.
* entropy ops         come from ``oracle/shim/compressai`` (restatement of the
                      un-vendored ``compressai==1.2.1``; PARITY UNPINNED).
* ``bits_from_likelihoods`` restates the inline bit estimate
                      ``video_model.py:39-42`` / ``:53-56``.
* ``pframe_hotpath``  composes one P-frame's worth of hot-path calls in the order
                      ``DeepSVC.forward`` issues them (``video_model.py:27-71``,
                      ``modules.py:148-170,424-438``, ``image_model.py:151-199``):
                      4 SpyNet-pyramid 3-ch warps, the 3-ch frame warp, the 64-ch
                      feature warp, 8+8 GaussianConditional slice calls with the
                      ``ste_round`` y_hat, 2 EntropyBottleneck calls, 4 bit sums.
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def _ensure_shim():
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)


_ensure_shim()
from compressai.entropy_models import EntropyBottleneck, GaussianConditional  # noqa: E402
from compressai.ops import LowerBound, ste_round  # noqa: E402

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """``image_model.py:18-19`` / ``modules.py:17-18``."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


_grid_cache = {}


def torch_warp(tensorInput: torch.Tensor, tensorFlow: torch.Tensor) -> torch.Tensor:
    """Backward bilinear warp, op-for-op as ``modules.py:25-62``."""
    key = (str(tensorInput.device), str(tensorFlow.size()))
    if key not in _grid_cache:
        B, _, H, W = tensorFlow.shape
        hor = torch.linspace(-1.0, 1.0, W).view(1, 1, 1, W).expand(B, -1, H, -1)
        ver = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1).expand(B, -1, -1, W)
        _grid_cache[key] = torch.cat([hor, ver], 1).to(tensorInput.device)
    flow = torch.cat([tensorFlow[:, 0:1, :, :] / ((tensorInput.size(3) - 1.0) / 2.0),
                      tensorFlow[:, 1:2, :, :] / ((tensorInput.size(2) - 1.0) / 2.0)], 1)
    grid = _grid_cache[key] + flow
    return F.grid_sample(input=tensorInput, grid=grid.permute(0, 2, 3, 1), mode="bilinear",
                         padding_mode="border", align_corners=True)


def spynet_level(im2: torch.Tensor, flow: torch.Tensor):
    """One SpyNet level's warp input (``modules.py:107-112,163-168``): (flow_up, warped)."""
    flow_up = F.interpolate(flow, (flow.size(2) * 2, flow.size(3) * 2), mode="bilinear",
                            align_corners=False) * 2.0
    return flow_up, torch_warp(im2, flow_up)


def warp_and_loss(ref_frame: torch.Tensor, flow: torch.Tensor, curr_frame: torch.Tensor):
    """``video_model.py:37-38``: (warped_frame, warp_loss)."""
    warped = torch_warp(ref_frame, flow)
    return warped, torch.mean((warped - curr_frame).pow(2))


def bits_from_likelihoods(likelihoods: torch.Tensor) -> torch.Tensor:
    """Total bits of one likelihood tensor: ``log(l).sum() / -ln 2``
    (``video_model.py:39-42`` before the division by the pixel count)."""
    return torch.log(likelihoods).sum() / (-math.log(2))


def make_entropy_models(ch_y: int, seed: int = 0, factor_std: float = 0.1):
    """(EntropyBottleneck(ch_y), GaussianConditional(None)) as constructed at
    ``image_model.py:148-149``, with the bottleneck's ``_factor`` parameters
    perturbed so that the tanh gate is exercised."""
    g = torch.Generator().manual_seed(seed)
    eb = EntropyBottleneck(ch_y)
    with torch.no_grad():
        for i in range(4):
            f = getattr(eb, f"_factor{i}")
            f.copy_(torch.randn(f.shape, generator=g) * factor_std)
        eb.quantiles[:, 0, 1] = torch.randn(ch_y, generator=g) * 0.3
    gc = GaussianConditional(None)
    gc.update_scale_table(get_scale_table())  # fills scale_table + CDF buffers
    return eb, gc


def codec_entropy_forward(eb, gc, y, z, scales, means, num_slices=8, training=False,
                          noise_y=None, noise_z=None):
    """The entropy part of ``ChannelSplitICIP2020ResB.forward``
    (``image_model.py:155-199``) with the conv transforms removed: the harness
    supplies per-slice ``scales`` / ``means`` directly.

    Returns (y_hat [ste-rounded about the mean], z_hat, y_likelihoods, z_likelihoods).
    ``noise_*``: explicit U(-1/2,1/2) draws for training mode (replaces the
    in-module ``uniform_`` so that both sides see identical noise).
    """
    if training and noise_z is not None:
        # EntropyBottleneck.forward in noise mode with an explicit draw
        perm = z.permute(1, 0, 2, 3).contiguous()
        vals = perm.reshape(perm.size(0), 1, -1) + noise_z.permute(1, 0, 2, 3).reshape(perm.size(0), 1, -1)
        lik = eb.likelihood_lower_bound(eb._likelihood(vals))
        z_lik = lik.reshape(perm.shape).permute(1, 0, 2, 3).contiguous()
    else:
        _, z_lik = eb(z, training=training)
    z_off = eb._get_medians()
    # image_model.py:160-162 (medians broadcast as [C,1,1] against [B,C,h,w])
    z_hat = ste_round(z - z_off) + z_off
    y_hat_slices, y_lik = [], []
    ys = y.chunk(num_slices, 1)
    ss = scales.chunk(num_slices, 1)
    ms = means.chunk(num_slices, 1)
    ns = noise_y.chunk(num_slices, 1) if noise_y is not None else [None] * num_slices
    for y_s, s_s, m_s, n_s in zip(ys, ss, ms, ns):
        if training and n_s is not None:
            out = y_s + n_s
            lik = gc.likelihood_lower_bound(gc._likelihood(out, s_s, m_s))
        else:
            _, lik = gc(y_s, s_s, m_s, training=training)
        y_lik.append(lik)
        y_hat_slices.append(ste_round(y_s - m_s) + m_s)  # image_model.py:183
    return torch.cat(y_hat_slices, 1), z_hat, torch.cat(y_lik, 1), z_lik


def pframe_hotpath(inputs: dict, models: dict, training: bool = False) -> dict:
    """One P-frame of warp+entropy work (no conv transforms), reference arithmetic.

    ``inputs`` comes from ``deepsvc_b200.synthetic.make_pframe_inputs`` (plain
    tensors); ``models`` maps "mv"/"res" to (EntropyBottleneck, GaussianConditional).
    """
    out = {}
    out["spynet"] = [torch_warp(im, fl) for im, fl in zip(inputs["pyr_img"], inputs["pyr_flow"])]
    out["warped_frame"] = torch_warp(inputs["ref_frame"], inputs["flow"])
    out["warped_feature"] = torch_warp(inputs["feature"], inputs["flow"])
    B, _, H, W = inputs["ref_frame"].shape
    pixels = B * H * W
    for name in ("mv", "res"):
        eb, gc = models[name]
        y_hat, z_hat, y_lik, z_lik = codec_entropy_forward(
            eb, gc, inputs[f"{name}_y"], inputs[f"{name}_z"], inputs[f"{name}_scales"],
            inputs[f"{name}_means"], training=training,
            noise_y=inputs.get(f"{name}_noise_y"), noise_z=inputs.get(f"{name}_noise_z"))
        out[f"{name}_y_hat"] = y_hat
        out[f"{name}_z_hat"] = z_hat
        out[f"{name}_y_lik"] = y_lik
        out[f"{name}_z_lik"] = z_lik
        bits = bits_from_likelihoods(y_lik) + bits_from_likelihoods(z_lik)
        out[f"bpp_{name}"] = bits / pixels
    out["bpp"] = out["bpp_mv"] + out["bpp_res"]
    return out


__all__ = ["torch_warp", "spynet_level", "warp_and_loss", "bits_from_likelihoods", "make_entropy_models", "codec_entropy_forward",
           "pframe_hotpath", "get_scale_table", "EntropyBottleneck", "GaussianConditional",
           "LowerBound", "ste_round"]
