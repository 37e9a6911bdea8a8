"""Quantised CDF tables for the range coder ("next" row f-2): the ``update()`` methods of
compressai 1.2.1's ``GaussianConditional`` / ``EntropyBottleneck`` as called through
``image_model.py:319-324`` and ``test_video.py:235``.  Run once per model load on
table-sized tensors (64 x ~3000).  The pmf is evaluated with the same torch ops ON THE SAME
DEVICE as compressai does (the parameters' / scale table's device: CPU and CUDA ``erfc`` /
``sigmoid`` differ in the last bit, and a table entry off by one makes bit streams that only
the same side can decode), then quantised on the host by
``csrc/coder.cpp::dsvc_pmf_to_quantized_cdf_host``."""
import numpy as np
import scipy.stats
import torch

from .ans import pmf_to_quantized_cdf


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length, precision):
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
    for i, p in enumerate(pmf):
        prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
        c = torch.IntTensor(pmf_to_quantized_cdf(prob.tolist(), precision))
        cdf[i, : c.size(0)] = c
    return cdf


def _std_cumulative(x):
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * x)


def gaussian_cdf_tables(scale_table: torch.Tensor, tail_mass: float, precision: int = 16):
    """(quantized_cdf [n, L+2] int32, offset [n] int32, cdf_length [n] int32), returned on the
    CPU; the pmf is evaluated on ``scale_table``'s device like ``GaussianConditional.update``."""
    scale_table = scale_table.detach().float()
    device = scale_table.device
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    pmf_center = torch.ceil(scale_table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = torch.max(pmf_length).item()
    samples = torch.abs(torch.arange(max_length, device=device).int() - pmf_center[:, None]).float()
    samples_scale = scale_table.unsqueeze(1).float()
    upper = _std_cumulative((0.5 - samples) / samples_scale)
    lower = _std_cumulative((-0.5 - samples) / samples_scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    cdf = _pmf_to_cdf(pmf.cpu(), tail.cpu(), pmf_length.cpu(), max_length, precision)
    return cdf, (-pmf_center).cpu(), (pmf_length + 2).cpu()


def bottleneck_cdf_tables(eb):
    """Tables of a (drop-in) EntropyBottleneck from its quantiles and CDF network, evaluated on
    the parameters' device with the module's own eager ``_logits_cumulative`` (the op sequence
    of ``EntropyBottleneck.update``); returned on the CPU."""
    with torch.no_grad():
        q = eb.quantiles.detach().float()
        device = q.device
        medians = q[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
        offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = pmf_length.max().item()
        samples = torch.arange(max_length, device=device)[None, :] + pmf_start[:, None, None]
        lower = eb._logits_cumulative(samples - 0.5, stop_gradient=True)
        upper = eb._logits_cumulative(samples + 0.5, stop_gradient=True)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        cdf = _pmf_to_cdf(pmf.cpu(), tail.cpu(), pmf_length.cpu(), max_length, eb.entropy_coder_precision)
    return cdf, offset.cpu(), (pmf_length + 2).cpu()
