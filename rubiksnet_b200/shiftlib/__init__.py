"""rubiksnet.shiftlib mirror.  Like the reference (rubiksnet/shiftlib/__init__.py:1-2) the package
namespace exports the modules; functionals live in shiftlib.rubiks3d.primitive / shiftlib.rubiks2d.primitive
(`shiftlib.rubiks2d` is the sub-package, `shiftlib.rubiks2d.rubiks2d` the functional)."""
from . import rubiks2d, rubiks3d  # noqa: F401
from .rubiks2d.layer import RubiksShift2D
from .rubiks2d.primitive import rubiks2d_backward, rubiks2d_forward  # noqa: F401
from .rubiks3d.layer import RubiksShift3D, RubiksShiftBase, create_3d_from_2d, init_shift1d_nfold, init_shift_uniform  # noqa: F401
from .rubiks3d.primitive import (RubiksShift3DFunc, RubiksShift3DFunction, rubiks_shift_3d,  # noqa: F401
                                 rubiks_shift_3d_backward, rubiks_shift_3d_forward)

__all__ = ["RubiksShift2D", "RubiksShift3D", "RubiksShiftBase"]
