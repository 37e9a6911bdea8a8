"""The C-ABI library loads and exports every symbol include/deepsvc_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "deepsvc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dsvc_[a-z0-9_]+)\s*\(", hdr)))


def test_build_and_load():
    import __graft_entry__ as g
    g.build()
    from deepsvc_b200 import _lib
    lib = _lib.load()
    assert lib.dsvc_abi_version() == 1


def test_every_declared_symbol_is_exported_and_bound():
    from deepsvc_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes table and header disagree"


def test_argument_validation_without_gpu():
    """Null pointers / bad shapes are rejected before any CUDA call."""
    from deepsvc_b200 import _lib
    lib = _lib.load()
    assert lib.dsvc_warp_fwd_f32(None, None, None, 1, 3, 8, 8, None, None, 1.0, 1.0, 1.0, 1.0,
                                 0, 0, 0, None, 0, None) == 1
    assert lib.dsvc_warp_workspace_bytes(1, 1088, 1920) >= 16  # scheduler state
    assert lib.dsvc_gc_fwd_f32(None, None, None, None, None, None, None, None, None, None, 0,
                               None, 0.11, 1e-9, 1, 16, 16, 16, 16, 16, None) == 1
    assert lib.dsvc_reduce_slots(1, 65280) == (65280 + 511) // 512
    assert lib.dsvc_reduce_slots(8, 2048) == 8 * 4
    assert lib.dsvc_eb_reduce_slots(1, 64, 510) == 64 * 4
    assert b"invalid argument" in lib.dsvc_error_string(1)


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import deepsvc_b200 as d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.torch_warp(torch.zeros(1, 3, 8, 8), torch.zeros(1, 2, 8, 8))
    gc = d.GaussianConditional(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gc(torch.zeros(1, 8, 4, 4), torch.ones(1, 8, 4, 4), torch.zeros(1, 8, 4, 4))
    with pytest.raises(ValueError, match="Invalid quantization mode"):
        gc.quantize(torch.zeros(1), "bogus")
    eb = d.EntropyBottleneck(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        eb(torch.zeros(1, 8, 4, 4))


def test_state_dict_names_match_reference_checkpoints(oracle):
    """Buffer / parameter names are what the reference's checkpoints contain
    (image_model.py:304-317, utils.py:114-123)."""
    import deepsvc_b200 as d
    assert sorted(d.GaussianConditional(None).state_dict()) == sorted(oracle.GaussianConditional(None).state_dict())
    assert sorted(d.EntropyBottleneck(6).state_dict()) == sorted(oracle.EntropyBottleneck(6).state_dict())


def test_torch_operator_library_loads_and_registers_ops():
    """libdeepsvc_b200_torch.so (csrc_torch/ops.cpp) loads without a GPU and registers the
    deepsvc_b200:: operators with the dispatcher (no compute calls here)."""
    import torch
    from deepsvc_b200 import _lib
    ops = _lib.torch_ops()
    assert ops is not None, "build() compiles deepsvc_b200/lib/libdeepsvc_b200_torch.so"
    assert ops.abi_version() == 1
    for name in ("warp_fwd", "warp_bwd", "torch_warp", "gc_fwd", "gaussian_conditional", "eb_fwd",
                 "entropy_bottleneck"):
        assert hasattr(torch.ops.deepsvc_b200, name)
        assert getattr(torch.ops.deepsvc_b200, name).default._schema is not None
