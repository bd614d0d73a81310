"""RubiksShift3D / RubiksShiftBase modules and shift initialisers -- mirror of
rubiksnet/shiftlib/rubiks3d/layer.py (reference)."""
import torch
import torch.nn as nn

from ...utils import make_tuple
from .primitive import rubiks_shift_3d

__all__ = ["RubiksShift3D", "RubiksShiftBase", "init_shift_uniform", "init_shift1d_nfold", "create_3d_from_2d"]


def init_shift_uniform(shift):
    """U(-1, 1) (layer.py:21-22)."""
    nn.init.uniform_(shift, -1, 1)


def init_shift1d_nfold(shift, nfold=8, noise=1e-3):
    """TSM-like init for a 1D shift (layer.py:25-40): first 1/nfold channels +1, next 1/nfold -1,
    the rest U(-noise, noise) so they stay off exact zero."""
    dim, channels = shift.size()
    assert dim == 1, "only works with rubiks1d"
    fold = channels // nfold
    with torch.no_grad():
        shift[:, :fold] = 1
        shift[:, fold:2 * fold] = -1
        shift[:, 2 * fold:].uniform_(-noise, noise)
    return shift


class RubiksShiftBase(nn.Module):
    """Holds the learnable `shift` [dim, C / shift_groups] and applies `shift_function` (layer.py:43-83)."""

    def __init__(self, num_channels, stride=1, padding=0, normalize_grad=True, normalize_t_factor=1.0,
                 shift_groups=1, quantize=False, *, dim, shift_function):
        super().__init__()
        assert num_channels % shift_groups == 0, "Does not satisfy num_channels % shift_groups == 0"
        self.num_channels = num_channels
        self.stride, self.padding = stride, padding
        self.normalize_grad, self.normalize_t_factor = normalize_grad, normalize_t_factor
        self.quantize = quantize
        self.shift = nn.Parameter(torch.zeros(dim, num_channels // shift_groups))
        init_shift_uniform(self.shift)
        self.shift_function = shift_function

    def forward(self, x):
        return self.shift_function(x, self.shift, stride=self.stride, padding=self.padding,
                                   normalize_grad=self.normalize_grad,
                                   normalize_t_factor=self.normalize_t_factor, quantize=self.quantize)

    def extra_repr(self):
        return "shift_channels={}".format(self.num_channels)


class RubiksShift3D(RubiksShiftBase):
    """Learnable (dt, dh, dw) shift per channel over [N,T,C,H,W] (layer.py:86-107)."""

    def __init__(self, num_channels, stride=(1, 1, 1), padding=(0, 0, 0), normalize_grad=True,
                 normalize_t_factor=1.0, quantize=False, shift_groups=1):
        super().__init__(num_channels, stride, padding, normalize_grad, normalize_t_factor, shift_groups,
                         quantize=quantize, dim=3, shift_function=rubiks_shift_3d)


def create_3d_from_2d(module_2d, init_mode="tsm", normalize_t_factor=1.0, quantize=False):
    """Builds a RubiksShift3D whose (H,W) rows copy a RubiksShift2D and whose T row follows
    `init_mode` in {"tsm", "tsm-g<std>", "uni<magnitude>", "none"} (layer.py:110-154)."""
    from ..rubiks2d.layer import RubiksShift2D

    assert isinstance(module_2d, RubiksShift2D)
    module_3d = RubiksShift3D(module_2d.num_channels, stride=(1, *make_tuple(module_2d.stride, 2)),
                              padding=(0, *make_tuple(module_2d.padding, 2)), normalize_grad=True,
                              normalize_t_factor=normalize_t_factor, quantize=quantize)
    with torch.no_grad():
        dims, channels = module_3d.shift.size()
        assert dims == 3, "INTERNAL ERROR"
        module_3d.shift[1:, :] = module_2d.shift
        t_row = module_3d.shift[0]
        fold = channels // 8
        if init_mode.startswith("tsm-g"):
            std = float(init_mode[5:]) or 1e-2
            t_row[:fold] = 1.0 + torch.randn(fold) * std
            t_row[fold:2 * fold] = -1.0 + torch.randn(fold) * std
            t_row[2 * fold:] = torch.randn(channels - 2 * fold) * std
        elif init_mode == "tsm":
            t_row[:fold] = 1
            t_row[fold:2 * fold] = -1
            t_row[2 * fold:] = 0
        elif init_mode.startswith("uni"):
            magnitude = float(init_mode[3:])
            assert magnitude > 0, f"uniform random magnitude must > 0: {magnitude}"
            t_row *= magnitude  # the constructor already drew U(-1, 1)
        elif init_mode.lower() == "none":
            module_3d.shift.fill_(float("nan"))  # must be loaded from a checkpoint downstream
        else:
            raise NotImplementedError(f"unknown init mode {init_mode}")
    return module_3d
