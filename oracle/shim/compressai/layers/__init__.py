"""Oracle restatement of the two compressai 1.2.1 layer helpers the reference
imports (``image_model.py:5``): plain 3x3 conv and sub-pixel 3x3 up-conv."""
import torch.nn as nn


def conv3x3(in_ch: int, out_ch: int, stride: int = 1) -> nn.Module:
    return nn.Conv2d(in_ch, out_ch, kernel_size=3, stride=stride, padding=1)


def subpel_conv3x3(in_ch: int, out_ch: int, r: int = 1) -> nn.Sequential:
    return nn.Sequential(
        nn.Conv2d(in_ch, out_ch * r ** 2, kernel_size=3, padding=1), nn.PixelShuffle(r)
    )


__all__ = ["conv3x3", "subpel_conv3x3"]
