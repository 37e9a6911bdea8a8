"""TEST INFRASTRUCTURE -- the algebra of the cell-order warp backward, restated in numpy.

``deepsvc_b200/csrc/warp_bwd_cell.cu`` does not scatter ``grad_out`` through the four bilinear
taps (ATen ``grid_sampler_2d_backward``, reached from ``/root/reference/modules.py:58-62`` by
autograd, ``Learner.py:1343``).  It files every output pixel under the CELL (floor of its clamped
source coordinate) it samples and evaluates both gradients cell by cell.  This module states that
reformulation in a dozen numpy lines so that ``tests/test_cell_order_cpu.py`` can check the
identity against autograd of the reference's own function on the CPU, independently of any CUDA
code.  Only ``tests/`` may import it.
"""
import numpy as np


def cell_order_backward(grad_out: np.ndarray, inp: np.ndarray, flow: np.ndarray):
    """(grad_input, grad_flow) of ``torch_warp(inp, flow)`` for float64 arrays [B,C,H,W] / [B,2,H,W]."""
    B, C, H, W = inp.shape
    gin = np.zeros_like(inp)
    gflow = np.zeros_like(flow)
    sx, sy = (W - 1.0) / 2.0, (H - 1.0) / 2.0
    # the reference builds its base grid with fp32 torch.linspace (modules.py:47-52)
    import torch
    lin_x = torch.linspace(-1.0, 1.0, W).double().numpy()
    lin_y = torch.linspace(-1.0, 1.0, H).double().numpy()
    for b in range(B):
        # modules.py:47-57 + grid_sampler_unnormalize / clip_coordinates (align_corners, border)
        ux = ((lin_x[None, :] + flow[b, 0] / sx) + 1.0) * 0.5 * (W - 1)
        uy = ((lin_y[:, None] + flow[b, 1] / sy) + 1.0) * 0.5 * (H - 1)
        ix, iy = np.clip(ux, 0.0, W - 1.0), np.clip(uy, 0.0, H - 1.0)
        free_x = (ux > 0.0) & (ux < W - 1.0)   # clip_coordinates_set_grad: zero where clipped
        free_y = (uy > 0.0) & (uy < H - 1.0)
        X, Y = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)   # the pixel's cell
        wx, wy = ix - X, iy - Y
        cell = (Y * W + X).ravel()
        # zero-padded input: taps outside the image are dropped by the reference (safe_add / within_bounds)
        pad = np.zeros((C, H + 1, W + 1), dtype=inp.dtype)
        pad[:, :H, :W] = inp[b]
        for c in range(C):
            g = grad_out[b, c]
            # four sums per cell over its pixels
            s0 = np.bincount(cell, weights=g.ravel(), minlength=H * W).reshape(H, W)
            s_x = np.bincount(cell, weights=(g * wx).ravel(), minlength=H * W).reshape(H, W)
            s_y = np.bincount(cell, weights=(g * wy).ravel(), minlength=H * W).reshape(H, W)
            s_xy = np.bincount(cell, weights=(g * wx * wy).ravel(), minlength=H * W).reshape(H, W)
            se, sw, ne = s_xy, s_y - s_xy, s_x - s_xy
            nw = s0 - s_x - s_y + s_xy
            # grad_input[Y][X] = NW of its own cell + NE of the west cell + SW of the north cell + SE of the north-west cell
            out = nw.copy()
            out[:, 1:] += ne[:, :-1]
            out[1:, :] += sw[:-1, :]
            out[1:, 1:] += se[:-1, :-1]
            gin[b, c] = out
            # grad_flow: the pixel's four input taps are the corners of ITS cell
            v00, v01 = pad[c, Y, X], pad[c, Y, X + 1]
            v10, v11 = pad[c, Y + 1, X], pad[c, Y + 1, X + 1]
            dxt, dxb, dyl, dyr = v01 - v00, v11 - v10, v10 - v00, v11 - v01
            gflow[b, 0] += g * (dxt + wy * (dxb - dxt))
            gflow[b, 1] += g * (dyl + wx * (dyr - dyl))
        # d(source coordinate)/d(flow) = ((W-1)/2) / sx = 1 where the coordinate is not clipped
        gflow[b, 0] *= free_x
        gflow[b, 1] *= free_y
    return gin, gflow
