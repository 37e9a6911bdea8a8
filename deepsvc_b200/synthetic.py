"""Synthetic P-frame hot-path inputs of the shapes ``DeepSVC.forward`` produces
(``video_model.py:27-71``).  Plain torch on the CPU generator so
that the CPU oracle and the GPU path see identical bits; no reference or oracle code
is involved.

Shapes for a padded frame [B,3,H,W] (H, W multiples of 64, ``modules.py:76-89``):
  pyr_img[k]   [B,3,H/2^(3-k),W/2^(3-k)], pyr_flow[k] [B,2,...]   k = 0..3  (SpyNet, modules.py:155-168)
  ref_frame    [B,3,H,W]      flow [B,2,H,W]                       (video_model.py:37)
  feature      [B,64,H,W]                                          (modules.py:429)
  mv_y/scales/means [B,64,H/16,W/16],  mv_z [B,64,H/64,W/64]       (image_model.py:152-181, N=64)
  res_y/...         [B,96,H/16,W/16],  res_z [B,96,H/64,W/64]      (N=96)
"""
import math

import torch
import torch.nn.functional as F

SEED = 16  # the reference's default seed (utils.py:16)
MV_CH, RES_CH, FEAT_CH, NUM_SLICES = 64, 96, 64, 8


def get_scale_table(lo=0.11, hi=256.0, levels=64):
    """The reference's scale table (``image_model.py:18-25``: SCALES_MIN/MAX/LEVELS)."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def smooth_flow(B, H, W, gen, sigma=4.0, jitter=0.25, scale=1.0):
    """SpyNet-like flow: N(0, sigma^2) px on a 1/16 grid, bilinearly upsampled x16
    (``modules.py:107-120,163``), plus per-pixel jitter."""
    h, w = max(H // 16, 1), max(W // 16, 1)
    coarse = torch.randn(B, 2, h, w, generator=gen) * sigma
    flow = F.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    flow = flow + torch.randn(B, 2, H, W, generator=gen) * jitter
    return (flow * scale).contiguous()


def gentle_flow(B, H, W, gen):
    """A low-gradient flow (N(0, 1) px nodes on the 1/16 grid = +-0.09 px/px, +-0.05 px jitter):
    not a SURVEY 8d family, used by one secondary bench line to show how much of the feature
    warp's shared-memory bank conflicts comes from the primary flow's +-0.35 px/px gradient."""
    return smooth_flow(B, H, W, gen, sigma=1.0, jitter=0.05)


def stress_flow(B, H, W, gen, sigma=16.0):
    return (torch.randn(B, 2, H, W, generator=gen) * sigma).contiguous()


def border_flow(B, H, W, gen, margin=64, reach=64.0):
    """Displacement pointing outward by up to `reach` px inside a `margin`-px frame."""
    ys = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(B, 1, H, W)
    xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(B, 1, H, W)
    u = torch.rand(B, 2, H, W, generator=gen) * reach
    fx = torch.where(xs < margin, -u[:, 0:1], torch.where(xs >= W - margin, u[:, 0:1], torch.zeros(())))
    fy = torch.where(ys < margin, -u[:, 1:2], torch.where(ys >= H - margin, u[:, 1:2], torch.zeros(())))
    return torch.cat([fx, fy], 1).contiguous()


def make_flow(kind, B, H, W, gen):
    """The three flow families of SURVEY 8d by name: smooth | stress | border."""
    if kind == "smooth":
        return smooth_flow(B, H, W, gen)
    if kind == "stress":
        return stress_flow(B, H, W, gen)
    if kind == "border":
        return border_flow(B, H, W, gen, margin=max(1, min(H, W) // 4), reach=40.0)
    if kind == "gentle":
        return gentle_flow(B, H, W, gen)
    raise ValueError(f"unknown flow kind {kind!r}")


def make_latents(B, C, h, w, gen, tie_frac=0.01, tail_frac=0.001):
    """mu ~ N(0,1); scale = exp(U(ln .05, ln 32)) (about 12 % below the 0.11 bound);
    y = mu + scale * N(0,1) with exact .5 ties and far-tail (likelihood floor) elements."""
    mu = torch.randn(B, C, h, w, generator=gen)
    lo, hi = math.log(0.05), math.log(32.0)
    scale = torch.exp(torch.rand(B, C, h, w, generator=gen) * (hi - lo) + lo)
    y = mu + scale * torch.randn(B, C, h, w, generator=gen)
    r = torch.rand(B, C, h, w, generator=gen)
    k = torch.randint(-3, 4, (B, C, h, w), generator=gen).float()
    y = torch.where(r < tie_frac, mu + k + 0.5, y)
    sign = torch.where(torch.rand(B, C, h, w, generator=gen) < 0.5, -1.0, 1.0)
    y = torch.where((r >= tie_frac) & (r < tie_frac + tail_frac), mu + sign * 40.0 * scale, y)
    return y.contiguous(), scale.contiguous(), mu.contiguous()


def make_pframe_inputs(B=1, H=256, W=448, seed=SEED, flow_kind="smooth", training=False,
                       feature_ch=FEAT_CH):
    """All tensors one P-frame's hot path consumes, on the CPU."""
    assert H % 64 == 0 and W % 64 == 0, "the reference pads frames to multiples of 64"
    gen = torch.Generator().manual_seed(seed)
    mk = {"smooth": smooth_flow, "stress": stress_flow, "border": border_flow, "gentle": gentle_flow}[flow_kind]
    d = {}
    d["pyr_img"], d["pyr_flow"] = [], []
    for k in range(4):
        s = 2 ** (3 - k)
        hh, ww = H // s, W // s
        d["pyr_img"].append(torch.rand(B, 3, hh, ww, generator=gen))
        if k == 0:  # SpyNet's coarsest level warps with the zero initial flow (modules.py:161-163)
            d["pyr_flow"].append(torch.zeros(B, 2, hh, ww))
        else:
            fl = mk(B, hh, ww, gen)
            d["pyr_flow"].append(fl / s if flow_kind == "smooth" else fl)
    d["ref_frame"] = torch.rand(B, 3, H, W, generator=gen)
    d["flow"] = mk(B, H, W, gen)
    d["feature"] = torch.randn(B, feature_ch, H, W, generator=gen)
    for name, C in (("mv", MV_CH), ("res", RES_CH)):
        y, s, m = make_latents(B, C, H // 16, W // 16, gen)
        d[f"{name}_y"], d[f"{name}_scales"], d[f"{name}_means"] = y, s, m
        d[f"{name}_z"] = torch.randn(B, C, H // 64, W // 64, generator=gen) * 3.0
        if training:
            d[f"{name}_noise_y"] = torch.rand(y.shape, generator=gen) - 0.5
            d[f"{name}_noise_z"] = torch.rand(d[f"{name}_z"].shape, generator=gen) - 0.5
    return d


def to_device(d, device):
    out = {}
    for k, v in d.items():
        out[k] = [t.to(device) for t in v] if isinstance(v, list) else v.to(device)
    return out


def pframe_algorithmic_bytes(B, H, W, coded=False, feature_ch=FEAT_CH):
    """Compulsory HBM traffic of one forward/estimate P-frame:
    warp = 4*B*h*w*(2C+2); gc (fused bits) = 16 B/elem; eb = 8 B/elem."""
    def warp(C, h, w):
        return 4 * B * h * w * (2 * C + 2)
    spynet = sum(warp(3, H >> k, W >> k) for k in range(4))
    frame = warp(3, H, W)
    feat = warp(feature_ch, H, W)
    ent = 16 * B * (MV_CH + RES_CH) * (H // 16) * (W // 16) + 8 * B * (MV_CH + RES_CH) * (H // 64) * (W // 64)
    return {"spynet": spynet, "frame": frame, "feature": feat, "entropy": ent,
            "total": spynet + frame + feat + ent}
