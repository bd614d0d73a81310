"""rubiksnet.shiftlib mirror: RubiksShift2D / RubiksShift3D modules and their functionals."""
from .rubiks2d import RubiksShift2D, VFS2DFunc, rubiks2d, rubiks2d_backward, rubiks2d_forward
from .rubiks3d import (RubiksShift3D, RubiksShift3DFunc, RubiksShift3DFunction, RubiksShiftBase,
                       create_3d_from_2d, init_shift1d_nfold, init_shift_uniform, rubiks_shift_3d,
                       rubiks_shift_3d_backward, rubiks_shift_3d_forward)

__all__ = [
    "RubiksShift2D", "RubiksShift3D", "RubiksShiftBase",
    "rubiks2d", "rubiks2d_forward", "rubiks2d_backward", "VFS2DFunc",
    "rubiks_shift_3d", "rubiks_shift_3d_forward", "rubiks_shift_3d_backward",
    "RubiksShift3DFunc", "RubiksShift3DFunction",
    "create_3d_from_2d", "init_shift_uniform", "init_shift1d_nfold",
]
