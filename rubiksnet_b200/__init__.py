"""rubiksnet_b200 -- B200-native (sm_100a) implementation of RubiksNet's learnable-shift hot path.

Public surface mirrors the reference package `rubiksnet`:
    rubiksnet_b200.RubiksNet(tier, num_classes, num_frames=8, variant="rubiks3d")
    rubiksnet_b200.shiftlib.{RubiksShift2D, RubiksShift3D, RubiksShiftBase, rubiks_shift_3d, rubiks2d, ...}
    rubiksnet_b200.attention_shift.AttentionShift
    rubiksnet_b200.rubiksnet_cuda  (the native module's six entry points, over the C ABI)
"""
from . import _lib, rubiksnet_cuda, shiftlib  # noqa: F401
from .attention_shift import AttentionShift  # noqa: F401
from .backbone import RubiksNetBackbone, RubiksShiftBlock  # noqa: F401
from .models import RubiksNet  # noqa: F401
from .shiftlib import RubiksShift2D, RubiksShift3D, RubiksShiftBase  # noqa: F401

__version__ = "0.1.0"
