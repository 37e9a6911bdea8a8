"""ctypes binding of the C ABI declared in ``include/deepsvc_b200.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (plain nvcc,
``-gencode arch=compute_100a,code=sm_100a``).  There is NO fallback: if the library
is missing, or a launcher reports an error, the calling op raises.
"""
import ctypes
import os
from ctypes import c_int, c_int32, c_int64, c_float, c_void_p, c_char_p, c_size_t

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdeepsvc_b200.so")
TORCH_LIB_PATH = os.path.join(_HERE, "lib", "libdeepsvc_b200_torch.so")

FLOW_MUL_RECIPROCAL = 0
FLOW_TRUE_DIVIDE = 1
WARP_AUTO, WARP_GATHER, WARP_TMA = 0, 1, 2
WARP_BWD_AUTO, WARP_BWD_DIRECT, WARP_BWD_STAGED, WARP_BWD_GATHER, WARP_BWD_CELL = 0, 1, 2, 3, 4
LAYOUT_NCHW, LAYOUT_NHWC = 0, 1
EB_PARAMS_PER_CHANNEL = 60

_P = c_void_p

# name -> (restype, argtypes); mirrors include/deepsvc_b200.h one to one
SIGNATURES = {
    "dsvc_abi_version": (c_int, []),
    "dsvc_error_string": (c_char_p, [c_int]),
    "dsvc_device_arch": (c_int, []),
    "dsvc_warp_fwd_f32": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                  c_float, c_float, c_float, c_float, c_int, c_int, c_int,
                                  _P, c_size_t, _P]),
    "dsvc_warp_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dsvc_warp_fwd2_f32": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P,
                                   c_float, c_float, c_float, c_float, c_int, _P, c_size_t, _P]),
    "dsvc_warp_bwd_f32": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                  c_float, c_float, c_float, c_float, c_int, c_int, _P]),
    "dsvc_warp_bwd_ws_f32": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                     c_float, c_float, c_float, c_float, c_int, c_int, _P, c_size_t, _P]),
    "dsvc_warp_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dsvc_warp_bwd_cell_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "dsvc_warp_fused_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P,
                                    c_float, c_float, c_float, c_float, c_int, _P]),
    "dsvc_warp_fused_slots": (c_int, [c_int, c_int, c_int]),
    "dsvc_blend_f32": (c_int, [_P, _P, _P, _P, c_int64, _P]),
    "dsvc_lrp_add_f32": (c_int, [_P, _P, _P, c_int64, _P]),
    "dsvc_lrp_add_bwd_f32": (c_int, [_P, _P, _P, c_int64, _P]),
    "dsvc_set_warp_bwd_algo": (c_int, [c_int]),
    "dsvc_reduce_slots": (c_int, [c_int64, c_int64]),
    "dsvc_gc_fwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P,
                                c_float, c_float, c_int64, c_int64,
                                c_int64, c_int64, c_int64, c_int64, _P]),
    "dsvc_gc_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_float, c_float,
                                c_int64, c_int64, c_int64, c_int64, c_int64, c_int64, _P]),
    "dsvc_eb_reduce_slots": (c_int, [c_int, c_int, c_int]),
    "dsvc_eb_pack_f32": (c_int, [_P, _P, c_int, _P]),
    "dsvc_eb_pack_bwd_f32": (c_int, [_P, _P, _P, c_int, _P]),
    "dsvc_eb_fwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "dsvc_eb_bwd_f32": (c_int, [_P, _P, _P, _P, _P, _P, c_float, c_int, c_int, c_int, _P]),
    "dsvc_bits_finalize_f64": (c_int, [_P, _P, _P, _P, c_int, _P]),
    # host-side range coder / CDF quantiser
    "dsvc_pmf_to_quantized_cdf_host": (c_int, [_P, c_int, c_int, _P]),
    "dsvc_rans_encoder_create": (c_void_p, []),
    "dsvc_rans_encoder_destroy": (None, [_P]),
    "dsvc_rans_encoder_push": (c_int, [_P, _P, _P, c_int64, _P, c_int, c_int, _P, _P]),
    "dsvc_rans_encoder_bound": (c_int64, [_P]),
    "dsvc_rans_encoder_flush": (c_int, [_P, _P, c_int64, _P]),
    "dsvc_rans_decoder_create": (c_void_p, [_P, c_int64]),
    "dsvc_rans_decoder_destroy": (None, [_P]),
    "dsvc_rans_decoder_decode": (c_int, [_P, _P, c_int64, _P, c_int, c_int, _P, _P, _P]),
    "dsvc_rans_encode_many": (c_int, [_P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int]),
    "dsvc_rans_decode_many": (c_int, [_P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_int]),
}

_lib = None


class DeepSVCNativeError(RuntimeError):
    pass


def load():
    """Load the native library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DeepSVCNativeError(
            f"deepsvc_b200 native library not found at {LIB_PATH}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.dsvc_abi_version() != 1:
        raise DeepSVCNativeError("deepsvc_b200 ABI version mismatch; rebuild the library")
    _lib = lib
    return lib


_ops = None
_ops_tried = False


def torch_ops():
    """``torch.ops.deepsvc_b200`` -- the TORCH_LIBRARY operator layer (``csrc_torch/ops.cpp``: C++
    autograd and output allocation over the same C ABI), or None when that library has not
    been built.  The eager drop-in entry points use it when present (one dispatcher hop per
    call); without it they go through ctypes.  Either way the arithmetic is the sm_100a kernels
    of the C-ABI library: set DSVC_NO_TORCH_OPS=1 to force the ctypes path."""
    global _ops, _ops_tried
    if _ops_tried:
        return _ops
    _ops_tried = True
    if os.environ.get("DSVC_NO_TORCH_OPS") == "1" or not os.path.isfile(TORCH_LIB_PATH):
        return None
    import torch
    load()                                  # the kernel library first (also checks the ABI version)
    torch.ops.load_library(TORCH_LIB_PATH)
    ops = torch.ops.deepsvc_b200
    if ops.abi_version() != 1:
        raise DeepSVCNativeError("deepsvc_b200 torch operator library: ABI version mismatch; rebuild")
    _ops = ops
    return ops


def check(err: int, what: str):
    if err != 0:
        msg = load().dsvc_error_string(err)
        raise DeepSVCNativeError(f"{what} failed: CUDA error {err} "
                                 f"({msg.decode() if msg else 'unknown'})")


def ptr(t):
    """Device pointer of a tensor (or NULL for None)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream
