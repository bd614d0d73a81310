"""RubiksShift2D module -- mirror of rubiksnet/shiftlib/rubiks2d/layer.py (reference)."""
import torch
import torch.nn as nn

from .primitive import rubiks2d

__all__ = ["RubiksShift2D", "init_shift_group"]


def init_shift_group(shift, kernel_size):
    """Zero-FLOP grouped-shift init: channels enumerate the K x K integer offsets (layer.py:6-15)."""
    k = kernel_size
    channels = shift.size(1)
    offsets = torch.arange(-(k // 2), k // 2 + 1, dtype=shift.dtype)
    groups = channels // k ** 2
    shift[0, :] = offsets.repeat(k * groups)
    shift[1, :] = offsets.repeat_interleave(k).repeat(groups)


class RubiksShift2D(nn.Module):
    """Learnable (dh, dw) shift per channel over [N,C,H,W] (layer.py:18-52)."""

    def __init__(self, num_channels, stride=1, padding=0, normalize_grad=True, quantize=False,
                 init_shift="uniform"):
        super().__init__()
        self.num_channels = num_channels
        self.stride, self.padding = stride, padding
        self.normalize_grad, self.quantize = normalize_grad, quantize
        self.shift = nn.Parameter(torch.zeros(2, num_channels))
        with torch.no_grad():
            if init_shift == "uniform":
                nn.init.uniform_(self.shift, -1, 1)
            elif init_shift.startswith("group"):
                group_kernel = int(init_shift[5:])
                assert group_kernel > 1
                init_shift_group(self.shift, group_kernel)
            else:
                raise NotImplementedError(f"unrecognized init shift {init_shift}")

    def forward(self, x):
        return rubiks2d(x, self.shift, stride=self.stride, padding=self.padding,
                        normalize_grad=self.normalize_grad, enable_shift_grad=True, quantize=self.quantize)

    def extra_repr(self):
        return "shift_channels={}".format(self.num_channels)
