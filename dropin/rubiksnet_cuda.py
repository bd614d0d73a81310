"""Top-level stand-in for the reference's native extension module.

The reference package imports its CUDA extension by this exact name (`import rubiksnet_cuda`,
rubiksnet/shiftlib/rubiks3d/primitive.py:2 and rubiks2d/primitive.py:2; pybind exports cuda_src/rubiks.cpp:384-396).
Putting this directory in front of sys.path (or on PYTHONPATH) makes the UNMODIFIED reference `rubiksnet` package run on
librubiks_b200.so through the C ABI of include/rubiks_b200.h:

    PYTHONPATH=/path/to/repo/dropin:/path/to/repo python scripts/test_installation.py

tests/test_gpu_dropin.py does exactly that with the reference's own scripts/test_installation.py recipe.
"""
from rubiksnet_b200.rubiksnet_cuda import *  # noqa: F401,F403
from rubiksnet_b200.rubiksnet_cuda import __all__  # noqa: F401

NATIVE_IMPL = "librubiks_b200"
