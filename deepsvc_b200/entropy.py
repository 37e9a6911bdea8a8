"""Drop-in replacements for the compressai 1.2.1 entropy-model classes the reference
uses (``image_model.py:4,7,148-149``): ``GaussianConditional``, ``EntropyBottleneck``,
``LowerBound``, ``ste_round``.

Same constructor arguments, method names, argument meaning, buffer / parameter names
(so the reference's checkpoints and ``update_registered_buffers`` calls at
``image_model.py:304-317`` keep working) and error behaviour (``ValueError`` on a bad
quantisation mode).  Every arithmetic method is ONE fused sm_100a launch from
``csrc/entropy.cu`` instead of the eager chains listed in SURVEY.md section 2.1.

Beyond the reference API each class offers ``forward_fused`` which also returns the
``ste_round`` y_hat (``image_model.py:183``) and per-CTA partial sums of ln(likelihood)
(the bit estimate of ``video_model.py:39-42``) from the same launch, so that a frame's
bits need no likelihood tensor at all.

No CPU path: non-CUDA tensors raise.
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib


# --------------------------------------------------------------------------- small ops
class _SteRoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, g):
        return g


def ste_round(x: Tensor) -> Tensor:
    """``compressai.ops.ste_round``: round(x) with identity gradient.  (Value-identical
    to ``round(x) - x.detach() + x``: both subtractions are exact in fp32.)"""
    if not (x.requires_grad and torch.is_grad_enabled()):
        return torch.round(x)
    return _SteRoundFn.apply(x)


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        return ((x >= bound) | (g < 0)) * g, None


class LowerBound(nn.Module):
    """``compressai.ops.LowerBound``: max(x, bound); gradient passes when x >= bound or
    when it would push x upward.  Inside the fused kernels the rule is applied in
    registers; this module exists for API / state_dict compatibility."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


# --------------------------------------------------------------------------- helpers
def _require_cuda_f32(name, *tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"deepsvc_b200.{name}: CUDA tensors required (no CPU fallback)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"deepsvc_b200.{name}: fp32 tensors required, got {t.dtype}")


STRICT_STRIDES = False  # True: raise on a layout the kernels cannot read in place (benchmarking)


def _dense_rows(t: Tensor) -> bool:
    return t.is_contiguous() or (t.dim() >= 2 and t.size(0) > 0 and t[0].is_contiguous())


def _in_place_layout(t: Optional[Tensor], what: str) -> Optional[Tensor]:
    """`t` itself when the kernels can read it in place (dense, or dense per batch row: the
    memory shape of ``y.chunk(num_slices, 1)`` slices, image_model.py:164).  Any other view the
    reference accepts -- e.g. the spatial crops ``mu[:, :, :h, :w]`` of image_model.py:171,175
    on an unpadded input -- is copied once, like the stock eager ops would (STRICT_STRIDES
    turns the copy into an error)."""
    if t is None or _dense_rows(t):
        return t
    if STRICT_STRIDES:
        raise RuntimeError(f"deepsvc_b200.{what}: unsupported strides (tensor must be contiguous, "
                           "or contiguous per batch row) and STRICT_STRIDES is set")
    return t.contiguous()


def _rows_of(t: Tensor):
    """(rows, inner, row_stride) of a tensor that is dense, or dense per batch row."""
    n = t.numel()
    if t.is_contiguous():
        return 1, n, n
    if t.dim() >= 2 and t.size(0) > 0 and t[0].is_contiguous():
        inner = n // t.size(0)
        return t.size(0), inner, t.stride(0)
    raise RuntimeError("deepsvc_b200: unsupported strides (tensor must be contiguous, or "
                       "contiguous per batch row)")


def _common_rows(tensors):
    """Express all (same-shape) tensors with one (rows, inner) split."""
    shape = tensors[0].shape
    infos = []
    for t in tensors:
        if t.shape != shape:
            raise RuntimeError(f"deepsvc_b200: shape mismatch {tuple(t.shape)} vs {tuple(shape)}")
        infos.append(_rows_of(t))
    rows = max(i[0] for i in infos)
    n = tensors[0].numel()
    if rows == 1:
        return 1, n, [n] * len(tensors)
    inner = n // rows
    strides = [inner if r == 1 else rs for (r, _, rs) in infos]
    return rows, inner, strides


def gc_launch(x, scales, means=None, noise=None, *, want_outputs=False, want_likelihood=False,
              want_y_hat=False, want_symbols=False, want_indexes=False, want_bits=False,
              scale_table=None, scale_bound=0.11, lik_bound=1e-9):
    """One fused GaussianConditional launch (no autograd).  Returns a dict with the
    requested outputs: outputs, likelihood, y_hat, symbols, indexes, bits_partials."""
    ops = _lib.torch_ops()
    if ops is not None and not STRICT_STRIDES:
        want = (1 if want_outputs else 0) | (2 if want_likelihood else 0) | (4 if want_y_hat else 0) | \
            (8 if want_symbols else 0) | (16 if want_indexes else 0) | (32 if want_bits else 0)
        o = ops.gc_fwd(x, scales, means, noise, scale_table if want_indexes else None, float(scale_bound),
                       float(lik_bound), want)
        return {"outputs": o[0], "likelihood": o[1], "y_hat": o[2], "symbols": o[3], "indexes": o[4],
                "bits_partials": o[5]}
    x, scales, means, noise = (_in_place_layout(t, "gaussian_conditional") for t in (x, scales, means, noise))
    ins = [t for t in (x, scales, means, noise) if t is not None]
    _require_cuda_f32("gaussian_conditional", *ins)
    rows, inner, st = _common_rows(ins)
    it = iter(st)
    x_rs = next(it)
    s_rs = next(it)
    m_rs = next(it) if means is not None else 0
    n_rs = next(it) if noise is not None else 0
    dev = x.device
    shape = tuple(x.shape)

    def new(dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=dev)

    res = {}
    res["outputs"] = new() if want_outputs else None
    res["likelihood"] = new() if want_likelihood else None
    res["y_hat"] = new() if want_y_hat else None
    res["symbols"] = new(torch.int32) if want_symbols else None
    res["indexes"] = new(torch.int32) if want_indexes else None
    lib = _lib.load()
    res["bits_partials"] = (torch.empty(lib.dsvc_reduce_slots(rows, inner), dtype=torch.float64,
                                        device=dev) if want_bits else None)
    if want_indexes:
        if scale_table is None or scale_table.numel() < 1:
            raise ValueError("build_indexes needs a non-empty scale_table (call update_scale_table)")
        if scale_table.device != dev or scale_table.dtype != torch.float32:
            raise RuntimeError("deepsvc_b200: scale_table must be fp32 on the input's device")
        scale_table = scale_table.contiguous()
    if x.numel() == 0:
        return res
    with torch.cuda.device(dev):
        err = lib.dsvc_gc_fwd_f32(
            x.data_ptr(), scales.data_ptr(), _lib.ptr(means), _lib.ptr(noise),
            _lib.ptr(res["outputs"]), _lib.ptr(res["likelihood"]), _lib.ptr(res["y_hat"]),
            _lib.ptr(res["symbols"]), _lib.ptr(res["indexes"]),
            _lib.ptr(scale_table) if want_indexes else None,
            int(scale_table.numel()) if want_indexes else 0,
            _lib.ptr(res["bits_partials"]), float(scale_bound), float(lik_bound),
            rows, inner, x_rs, s_rs, m_rs, n_rs, _lib.stream_ptr(dev))
    _lib.check(err, "dsvc_gc_fwd_f32")
    return res


class _GaussianConditionalFn(torch.autograd.Function):
    """(outputs, likelihood, y_hat, bits_partials) with the reference's gradients:
    likelihood -> x, scales, means (LowerBound rules inside); outputs -> x (noise mode)
    or means (round mode); y_hat -> x (straight-through)."""

    @staticmethod
    def forward(ctx, x, scales, means, noise, scale_bound, lik_bound, want_bits):
        x, scales, means, noise = (_in_place_layout(t, "GaussianConditional") for t in (x, scales, means, noise))
        r = gc_launch(x, scales, means, noise, want_outputs=True, want_likelihood=True,
                      want_y_hat=True, want_bits=want_bits, scale_bound=scale_bound,
                      lik_bound=lik_bound)
        ctx.save_for_backward(x, scales, means, noise)
        ctx.bounds = (scale_bound, lik_bound)
        ctx.set_materialize_grads(False)
        bits = r["bits_partials"]
        if bits is None:
            bits = torch.empty(0, dtype=torch.float64, device=x.device)
        ctx.mark_non_differentiable(bits)
        return r["outputs"], r["likelihood"], r["y_hat"], bits

    @staticmethod
    def backward(ctx, g_out, g_lik, g_yhat, _g_bits):
        x, scales, means, noise = ctx.saved_tensors
        scale_bound, lik_bound = ctx.bounds
        need_x, need_s, need_m = ctx.needs_input_grad[:3]
        need_m = need_m and means is not None
        gx = gs = gm = None
        if g_lik is not None and (need_x or need_s or need_m):
            g_lik = g_lik.contiguous()
            ins = [t for t in (x, scales, means, noise) if t is not None]
            rows, inner, st = _common_rows(ins)
            it = iter(st)
            x_rs, s_rs = next(it), next(it)
            m_rs = next(it) if means is not None else 0
            n_rs = next(it) if noise is not None else 0
            dev = x.device

            def new():
                return torch.empty(x.shape, dtype=torch.float32, device=dev)

            gx = new() if need_x else None
            gs = new() if need_s else None
            gm = new() if need_m else None
            if x.numel():
                lib = _lib.load()
                with torch.cuda.device(dev):
                    err = lib.dsvc_gc_bwd_f32(
                        g_lik.data_ptr(), x.data_ptr(), scales.data_ptr(), _lib.ptr(means),
                        _lib.ptr(noise), _lib.ptr(gx), _lib.ptr(gs), _lib.ptr(gm),
                        float(scale_bound), float(lik_bound), rows, inner, x_rs, s_rs, m_rs, n_rs,
                        _lib.stream_ptr(dev))
                _lib.check(err, "dsvc_gc_bwd_f32")

        def acc(a, b):
            return b if a is None else (a if b is None else a + b)

        if g_out is not None:
            if noise is not None:
                if need_x:
                    gx = acc(gx, g_out)
            elif need_m:
                gm = acc(gm, g_out)
        if g_yhat is not None and need_x:
            gx = acc(gx, g_yhat)
        return gx, gs, gm, None, None, None, None


# --------------------------------------------------------------------------- base class
class EntropyModel(nn.Module):
    """Mirror of ``compressai.entropy_models.EntropyModel`` (quantisation modes + CDF
    buffers).  ``entropy_coder`` arguments are accepted for signature compatibility."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        self._lik_bound = float(likelihood_bound) if self.use_likelihood_bound else float("-inf")
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._tables_gen = 0        # bumped whenever the CDF buffers are replaced
        self._tables_cache = None

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        _require_cuda_f32("quantize", inputs, means)
        if mode == "noise":
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)
            return inputs + noise
        if means is not None and means.shape != inputs.shape:
            means = means.expand_as(inputs).contiguous()
        # scales are irrelevant for pure quantisation: alias the input
        r = gc_launch(inputs, inputs, means, None, want_y_hat=(mode == "dequantize"),
                      want_symbols=(mode == "symbols"))
        return r["y_hat"] if mode == "dequantize" else r["symbols"]

    def _quantize(self, inputs, mode, means=None):
        return self.quantize(inputs, mode, means)

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None,
                   dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    @classmethod
    def _dequantize(cls, inputs, means=None):
        return cls.dequantize(inputs, means)

    # ---- range coding (compressai EntropyModel.compress / .decompress; call sites
    # image_model.py:206-207 through EntropyBottleneck).  Symbols are computed by the fused
    # kernel and leave the device in ONE int32 copy per batch item; the coding loop is
    # csrc/coder.cpp.
    def _check_tables(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")

    def _invalidate_tables(self):
        self._tables_gen += 1
        self._tables_cache = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._invalidate_tables()       # checkpoint CDFs replace the buffers' contents

    def __setattr__(self, name, value):
        # update_registered_buffers (image_model.py:304-317) and update() assign new buffers
        if name in ("_quantized_cdf", "_offset", "_cdf_length") and "_tables_gen" in self.__dict__:
            self.__dict__["_tables_gen"] += 1
            self.__dict__["_tables_cache"] = None
        super().__setattr__(name, value)

    def _cdf_tables(self):
        """Host copy of the coder tables; keyed on a generation counter (addresses can be
        recycled by the caching allocator) plus the buffers' in-place version counters."""
        from .ans import CdfTables
        key = (self._tables_gen, self._quantized_cdf._version, self._offset._version,
               self._cdf_length._version)
        if self._tables_cache is None or self._tables_cache[0] != key:
            self._tables_cache = (key, CdfTables(self._quantized_cdf, self._cdf_length, self._offset))
        return self._tables_cache[1]

    def compress(self, inputs, indexes, means=None):
        from .ans import RansEncoder
        symbols = self.quantize(inputs, "symbols", means)
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_tables()
        tables = self._cdf_tables()
        sym = symbols.cpu()
        idx = indexes.int().cpu()
        return [RansEncoder().encode_with_indexes(sym[i], idx[i], tables) for i in range(sym.size(0))]

    def decompress(self, strings, indexes, dtype: torch.dtype = torch.float, means=None):
        from .ans import RansDecoder
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        tables = self._cdf_tables()
        idx = indexes.int().cpu()
        out = torch.empty(indexes.size(), dtype=torch.int32)
        for i, s in enumerate(strings):
            dec = RansDecoder()
            dec.set_stream(s)
            out[i] = torch.from_numpy(dec.decode_stream_array(idx[i], tables)).reshape(out[i].size())
        dev = means.device if means is not None else self._quantized_cdf.device
        return self.dequantize(out.to(dev), means, dtype)


# --------------------------------------------------------------------------- Gaussian
class GaussianConditional(EntropyModel):
    """Drop-in for ``compressai.entropy_models.GaussianConditional`` as constructed at
    ``image_model.py:149`` (``GaussianConditional(None)``)."""

    def __init__(self, scale_table, *args, scale_bound: float = 0.11, tail_mass: float = 1e-9,
                 **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self._scale_bound = float(np.float32(scale_bound))
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table",
            self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    def _standardized_cumulative(self, inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        return half * torch.erfc(const * inputs)

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        from .cdf import gaussian_cdf_tables
        cdf, offset, length = gaussian_cdf_tables(self.scale_table, self.tail_mass,
                                                  self.entropy_coder_precision)
        dev = self.scale_table.device
        self._quantized_cdf = cdf.to(dev)
        self._offset = offset.to(dev)
        self._cdf_length = length.to(dev)

    def _bounds(self):
        return self._scale_bound, float(np.float32(self._lik_bound))

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """(outputs, likelihood) exactly as the reference call at ``image_model.py:181``."""
        out, lik, _, _ = self._run(inputs, scales, means, training, False)
        return out, lik

    def forward_fused(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                      training: Optional[bool] = None, noise: Optional[Tensor] = None):
        """One launch for ``image_model.py:181`` + ``:183`` + the log-sum of
        ``video_model.py:39-42``: returns (y_hat, likelihood, ln_lik_partials) where
        ``y_hat = ste_round(inputs - means) + means`` and ``ln_lik_partials.sum() /
        -ln 2`` is the slice's bit count."""
        _, lik, y_hat, bits = self._run(inputs, scales, means, training, True, noise)
        return y_hat, lik, bits

    def _run(self, inputs, scales, means, training, want_bits, noise=None):
        if training is None:
            training = self.training
        _require_cuda_f32("GaussianConditional", inputs, scales, means)
        if training and noise is None:
            half = float(0.5)
            noise = torch.empty_like(inputs).uniform_(-half, half)  # same draw as the reference
        if not training:
            noise = None
        sb, lb = self._bounds()
        ops = _lib.torch_ops()
        if ops is not None and not STRICT_STRIDES:
            return ops.gaussian_conditional(inputs, scales, means, noise, sb, lb, want_bits)
        return _GaussianConditionalFn.apply(inputs, scales, means, noise, sb, lb, want_bits)

    def likelihood_bits(self, inputs, scales, means=None):
        """Inference-only: y_hat and ln-likelihood partials without materialising the
        likelihood tensor (16 B/element)."""
        _require_cuda_f32("GaussianConditional", inputs, scales, means)
        sb, lb = self._bounds()
        r = gc_launch(inputs, scales, means, None, want_y_hat=True, want_bits=True,
                      scale_bound=sb, lik_bound=lb)
        return r["y_hat"], r["bits_partials"]

    def build_indexes(self, scales: Tensor) -> Tensor:
        """``image_model.py:237,286``: 6-step binary search instead of the 63-launch loop."""
        _require_cuda_f32("build_indexes", scales)
        sb, _ = self._bounds()
        r = gc_launch(scales, scales, None, None, want_indexes=True,
                      scale_table=self.scale_table, scale_bound=sb)
        return r["indexes"]

    def quantize_and_index(self, inputs, scales, means=None):
        """Codec path of ``image_model.py:237-239`` in one launch:
        (symbols int32, indexes int32, y_hat = symbols + means)."""
        _require_cuda_f32("quantize_and_index", inputs, scales, means)
        sb, _ = self._bounds()
        r = gc_launch(inputs, scales, means, None, want_symbols=True, want_indexes=True,
                      want_y_hat=True, scale_table=self.scale_table, scale_bound=sb)
        return r["symbols"], r["indexes"], r["y_hat"]


# --------------------------------------------------------------------------- bottleneck
def _raw15(eb):
    """The 15 raw tensors in the order of ``dsvc_eb_pack_f32``."""
    d = eb._parameters
    return ([d[f"_matrix{i}"] for i in range(5)] + [d[f"_bias{i}"] for i in range(5)] +
            [d[f"_factor{i}"] for i in range(4)] + [d["quantiles"]])


def _ptr15(tensors):
    import ctypes
    return (ctypes.c_void_p * 15)(*[None if t is None else t.data_ptr() for t in tensors])


class _PackFn(torch.autograd.Function):
    """raw parameters -> packed [C, 60] in one launch; backward in one more (the eager version is
    ~15 launches forward and ~25 backward per bottleneck: most of a training step's kernel nodes)."""

    @staticmethod
    def forward(ctx, *raw):
        raw = tuple(t.contiguous() for t in raw)
        C = raw[14].size(0)
        packed = torch.empty(C, _lib.EB_PARAMS_PER_CHANNEL, dtype=torch.float32, device=raw[0].device)
        with torch.cuda.device(raw[0].device):
            err = _lib.load().dsvc_eb_pack_f32(_ptr15(raw), packed.data_ptr(), C, _lib.stream_ptr(raw[0].device))
        _lib.check(err, "dsvc_eb_pack_f32")
        ctx.save_for_backward(*raw)
        return packed

    @staticmethod
    def backward(ctx, g):
        raw = ctx.saved_tensors
        g = g.contiguous()
        grads = [torch.empty_like(t) if ctx.needs_input_grad[i] else None for i, t in enumerate(raw)]
        with torch.cuda.device(g.device):
            err = _lib.load().dsvc_eb_pack_bwd_f32(_ptr15(raw), g.data_ptr(), _ptr15(grads), raw[14].size(0),
                                                   _lib.stream_ptr(g.device))
        _lib.check(err, "dsvc_eb_pack_bwd_f32")
        return tuple(grads)


def pack_bottleneck_params(eb: "EntropyBottleneck") -> Tensor:
    """[C, 60] fp32: softplus'd matrices, biases, tanh'd factors of the 1-3-3-3-3-1
    network (``EntropyBottleneck._logits_cumulative``), the median, one pad.  Differentiable:
    autograd reaches the raw parameters (one fused launch each way on CUDA; eager torch ops for
    a module that still lives on the CPU)."""
    if eb.filters != (3, 3, 3, 3):
        raise NotImplementedError("fused EntropyBottleneck supports filters=(3,3,3,3) "
                                  "(the reference's configuration)")
    C = eb.channels
    raw = _raw15(eb)
    if all(t.is_cuda and t.dtype == torch.float32 for t in raw):
        return _PackFn.apply(*raw)
    parts = []
    for i in range(5):
        parts.append(F.softplus(getattr(eb, f"_matrix{i}")).reshape(C, -1))
        parts.append(getattr(eb, f"_bias{i}").reshape(C, -1))
        if i < 4:
            parts.append(torch.tanh(getattr(eb, f"_factor{i}")).reshape(C, -1))
    parts.append(eb.quantiles[:, 0, 1:2])
    parts.append(torch.zeros(C, 1, dtype=torch.float32, device=eb.quantiles.device))
    packed = torch.cat(parts, 1)
    assert packed.shape[1] == _lib.EB_PARAMS_PER_CHANNEL
    return packed.contiguous()


def eb_launch(z, packed, noise=None, *, want_outputs=False, want_likelihood=False,
              want_z_hat=False, want_bits=False, lik_bound=1e-9):
    _require_cuda_f32("EntropyBottleneck", z, packed, noise)
    if not z.is_contiguous() or (noise is not None and not noise.is_contiguous()):
        raise RuntimeError("deepsvc_b200.EntropyBottleneck: contiguous [B,C,...] input required")
    B, C = z.shape[0], z.shape[1]
    S = z.numel() // max(B * C, 1)
    if packed.shape != (C, _lib.EB_PARAMS_PER_CHANNEL):
        raise RuntimeError("deepsvc_b200.EntropyBottleneck: channel mismatch")
    lib = _lib.load()
    res = {
        "outputs": torch.empty_like(z) if want_outputs else None,
        "likelihood": torch.empty_like(z) if want_likelihood else None,
        "z_hat": torch.empty_like(z) if want_z_hat else None,
        "bits_partials": (torch.empty(lib.dsvc_eb_reduce_slots(B, C, S), dtype=torch.float64,
                                      device=z.device) if want_bits else None),
    }
    if z.numel() == 0:
        return res
    with torch.cuda.device(z.device):
        err = lib.dsvc_eb_fwd_f32(z.data_ptr(), _lib.ptr(noise), packed.data_ptr(),
                                  _lib.ptr(res["outputs"]), _lib.ptr(res["likelihood"]),
                                  _lib.ptr(res["z_hat"]), _lib.ptr(res["bits_partials"]),
                                  float(np.float32(lik_bound)), B, C, S, _lib.stream_ptr(z.device))
    _lib.check(err, "dsvc_eb_fwd_f32")
    return res


class _EntropyBottleneckFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, packed, noise, lik_bound, want_bits):
        packed = packed.contiguous()
        r = eb_launch(z, packed, noise, want_outputs=True, want_likelihood=True, want_z_hat=True,
                      want_bits=want_bits, lik_bound=lik_bound)
        ctx.save_for_backward(z, packed, noise)
        ctx.lik_bound = lik_bound
        ctx.set_materialize_grads(False)
        bits = r["bits_partials"]
        if bits is None:
            bits = torch.empty(0, dtype=torch.float64, device=z.device)
        ctx.mark_non_differentiable(bits)
        return r["outputs"], r["likelihood"], r["z_hat"], bits

    @staticmethod
    def backward(ctx, g_out, g_lik, g_zhat, _g_bits):
        z, packed, noise = ctx.saved_tensors
        need_z, need_p = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gz = gp = None
        B, C = z.shape[0], z.shape[1]
        S = z.numel() // max(B * C, 1)
        if g_lik is not None and (need_z or need_p) and z.numel():
            g_lik = g_lik.contiguous()
            gz = torch.empty_like(z) if need_z else None
            gp = torch.zeros_like(packed) if need_p else None
            lib = _lib.load()
            with torch.cuda.device(z.device):
                err = lib.dsvc_eb_bwd_f32(g_lik.data_ptr(), z.data_ptr(), _lib.ptr(noise),
                                          packed.data_ptr(), _lib.ptr(gz), _lib.ptr(gp),
                                          float(np.float32(ctx.lik_bound)), B, C, S,
                                          _lib.stream_ptr(z.device))
            _lib.check(err, "dsvc_eb_bwd_f32")

        def acc(a, b):
            return b if a is None else (a if b is None else a + b)

        if g_out is not None:
            if noise is not None:
                if need_z:
                    gz = acc(gz, g_out)
            elif need_p:  # round mode: d outputs / d median = 1
                gm = torch.zeros_like(packed)
                gm[:, 58] = g_out.transpose(0, 1).reshape(C, -1).sum(1)
                gp = acc(gp, gm)
        if g_zhat is not None and need_z:
            gz = acc(gz, g_zhat)  # straight-through (image_model.py:160-162)
        return gz, gp, None, None, None


class EntropyBottleneck(EntropyModel):
    """Drop-in for ``compressai.entropy_models.EntropyBottleneck`` (``image_model.py:148``)."""

    _offset: Tensor

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))
        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed_cache = None

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _param_list(self):
        d = self._parameters   # (looked up per call: swap_entropy_models / load_state_dict may rebind entries)
        ps = [d["quantiles"]]
        for i in range(5):
            ps.append(d[f"_matrix{i}"])
            ps.append(d[f"_bias{i}"])
            if i < 4:
                ps.append(d[f"_factor{i}"])
        return ps

    def packed_params(self, differentiable: bool) -> Tensor:
        """Packed [C,60] parameters; cached (keyed on parameter versions) when no
        gradient is needed, so that inference is a single launch."""
        if differentiable:
            return pack_bottleneck_params(self)
        ps = self._param_list()
        # in-place updates (optimizer steps, load_state_dict) bump the version counters; a move of the
        # module (.to / .cuda) replaces every parameter's storage, the first one's included
        key = (ps[0].data_ptr(),) + tuple(p._version for p in ps)
        if self._packed_cache is None or self._packed_cache[0] != key:
            with torch.no_grad():
                self._packed_cache = (key, pack_bottleneck_params(self))
        return self._packed_cache[1]

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        # eager version, used only by loss()/update() on [C,1,3]-sized inputs
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(F.softplus(matrix), logits)
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def loss(self) -> Tensor:
        """Auxiliary quantile loss (``video_model.py:170-177``): host-sized ([C,1,3])."""
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        from .cdf import bottleneck_cdf_tables
        cdf, offset, length = bottleneck_cdf_tables(self)
        dev = self.quantiles.device
        self._quantized_cdf = cdf.to(dev)
        self._offset = offset.to(dev)
        self._cdf_length = length.to(dev)
        return True

    @staticmethod
    def _build_indexes(size):
        dims = len(size)
        N = size[0]
        C = size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(C).view(*view_dims)
        indexes = indexes.int()
        return indexes.repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        """``image_model.py:206``: per-channel table index, medians as means."""
        indexes = self._build_indexes(x.size())
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        """``image_model.py:207,260``."""
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians.dtype, medians)

    def _run(self, x, training, want_bits, noise=None):
        if training is None:
            training = self.training
        _require_cuda_f32("EntropyBottleneck", x)
        if x.dim() < 2 or x.size(1) != self.channels:
            raise RuntimeError("deepsvc_b200.EntropyBottleneck: expected [B, C, ...] input")
        if training and noise is None:
            half = float(0.5)
            # the reference draws the noise on the [C, 1, B*S] permuted view
            n = torch.empty(self.channels, 1, x.numel() // self.channels, device=x.device,
                            dtype=x.dtype).uniform_(-half, half)
            noise = n.reshape(self.channels, x.size(0), -1).transpose(0, 1).reshape(x.shape).contiguous()
        if not training:
            noise = None
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._parameters.values())
        packed = self.packed_params(need_grad)
        lb = float(np.float32(self._lik_bound))
        ops = _lib.torch_ops()
        if ops is not None:
            return ops.entropy_bottleneck(x, packed, noise, lb, want_bits)
        return _EntropyBottleneckFn.apply(x.contiguous() if not x.is_contiguous() else x,
                                          packed, noise, lb, want_bits)

    def forward(self, x: Tensor, training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """(outputs, likelihood) as the reference call at ``image_model.py:155``."""
        out, lik, _, _ = self._run(x, training, False)
        return out, lik

    def forward_fused(self, x: Tensor, training: Optional[bool] = None,
                      noise: Optional[Tensor] = None):
        """(z_hat, likelihood, ln_lik_partials): ``image_model.py:155`` + ``:160-162``."""
        _, lik, z_hat, bits = self._run(x, training, True, noise)
        return z_hat, lik, bits

    def likelihood_bits(self, x: Tensor):
        """Inference-only: (z_hat, ln-likelihood partials), no likelihood tensor."""
        _require_cuda_f32("EntropyBottleneck", x)
        r = eb_launch(x, self.packed_params(False), None, want_z_hat=True, want_bits=True,
                      lik_bound=self._lik_bound)
        return r["z_hat"], r["bits_partials"]


def bits_finalize(partials: Tensor, seg_offsets: Tensor, scales: Tensor, out: Optional[Tensor] = None):
    """out[i] = scales[i] * sum(partials[seg[i]:seg[i+1]]) in fp64, fixed order.  With
    scales = -1 / (ln 2 * pixels) this is the bpp expression of ``video_model.py:39-42``."""
    nseg = seg_offsets.numel() - 1
    if out is None:
        out = torch.empty(nseg, dtype=torch.float64, device=partials.device)
    lib = _lib.load()
    with torch.cuda.device(partials.device):
        err = lib.dsvc_bits_finalize_f64(partials.data_ptr(), seg_offsets.data_ptr(),
                                         scales.data_ptr(), out.data_ptr(), nseg,
                                         _lib.stream_ptr(partials.device))
    _lib.check(err, "dsvc_bits_finalize_f64")
    return out


def bpp_scale(pixels: int) -> float:
    return -1.0 / (math.log(2) * pixels)


__all__ = ["EntropyModel", "GaussianConditional", "EntropyBottleneck", "LowerBound", "ste_round",
           "gc_launch", "eb_launch", "pack_bottleneck_params", "bits_finalize", "bpp_scale"]
