"""Oracle restatement of compressai 1.2.1 ``ops/ops.py`` and ``ops/bound_ops.py``.

Call sites in the reference: ``image_model.py:7,162,183`` (ste_round);
LowerBound is used inside the entropy models (scale bound 0.11, likelihood
bound 1e-9).
"""
import torch
import torch.nn as nn
from torch import Tensor


def ste_round(x: Tensor) -> Tensor:
    """Rounding with a straight-through (identity) gradient."""
    return torch.round(x) - x.detach() + x


def lower_bound_fwd(x: Tensor, bound: Tensor) -> Tensor:
    return torch.max(x, bound)


def lower_bound_bwd(x: Tensor, bound: Tensor, grad_output: Tensor):
    pass_through_if = (x >= bound) | (grad_output < 0)
    return pass_through_if * grad_output, None


class LowerBoundFunction(torch.autograd.Function):
    """max(x, bound) whose gradient also passes when it pushes x upward."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return lower_bound_fwd(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        return lower_bound_bwd(x, bound, grad_output)


class LowerBound(nn.Module):
    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return LowerBoundFunction.apply(x, self.bound)


__all__ = ["ste_round", "LowerBound", "LowerBoundFunction"]
