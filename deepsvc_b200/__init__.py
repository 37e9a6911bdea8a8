"""deepsvc_b200 -- B200-native (sm_100a) warp + entropy P-frame hot path of DeepSVC.

Host-side mirror of the reference's operator interface for this path
(``modules.torch_warp``; compressai's ``GaussianConditional`` / ``EntropyBottleneck``
/ ``ste_round`` / ``LowerBound``) on top of the C ABI in ``include/deepsvc_b200.h``.
Importing the package does not require a GPU; calling an op does (no CPU fallback).
"""
from . import _lib, ans
from .warp import (torch_warp, warp_forward, warp_forward2, warp_backward, set_flow_arithmetic,
                   set_warp_algorithm)
from .entropy import (EntropyBottleneck, EntropyModel, GaussianConditional, LowerBound, ste_round,
                      bits_finalize, bpp_scale)
from .fused import lrp_add, mc_blend, spynet_level_warp, warp_with_mse
from .patch import patch_reference, swap_entropy_models, unpatch_reference

__version__ = "0.1.0"

__all__ = ["torch_warp", "warp_forward", "warp_forward2", "warp_backward", "set_flow_arithmetic",
           "set_warp_algorithm", "EntropyBottleneck", "EntropyModel", "GaussianConditional",
           "LowerBound", "ste_round", "bits_finalize", "bpp_scale", "patch_reference",
           "unpatch_reference", "swap_entropy_models", "spynet_level_warp", "warp_with_mse", "mc_blend", "lrp_add"]
