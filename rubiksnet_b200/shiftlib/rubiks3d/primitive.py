"""3D learnable shift: primitives, autograd function and functional.

Host-side mirror of rubiksnet/shiftlib/rubiks3d/primitive.py (reference): same public names,
keyword signatures, defaults, assertions and return conventions, over librubiks_b200's C ABI.
Extensions: float16 / bfloat16 activations (shift may stay float32), gradients that are not needed are
not computed, and launches follow torch's current stream.
"""
import torch
import torch.nn as nn

from ... import rubiksnet_cuda as _native
from ...utils import allocate_output, make_tuple

__all__ = [
    "rubiks_shift_3d_forward", "rubiks_shift_3d_backward", "rubiks_shift_3d",
    "RubiksShift3DFunc", "RubiksShift3DFunction", "compute_output_shape",
]

_FLOAT_TYPES = (torch.float32, torch.float64, torch.float16, torch.bfloat16)


def _out_len(size, stride, pad):
    # (in + 2p - 1) / s + 1 with C++ integer division (cuda_src/rubiks.cpp:166); the reference's Python
    # side computes the same in float and truncates (primitive.py:24-25,47)
    return (size + 2 * pad - 1) // stride + 1


def compute_output_shape(x, stride, padding, shift_dim=3):
    """Output size of a shift over [N,T,C,H,W] (primitive.py:28-47)."""
    n, t, c, h, w = x.size()
    s, p = make_tuple(stride, shift_dim), make_tuple(padding, shift_dim)
    if shift_dim == 1:
        t = _out_len(t, s[0], p[0])
    elif shift_dim == 2:
        h, w = _out_len(h, s[0], p[0]), _out_len(w, s[1], p[1])
    elif shift_dim == 3:
        t, h, w = _out_len(t, s[0], p[0]), _out_len(h, s[1], p[1]), _out_len(w, s[2], p[2])
    else:
        raise NotImplementedError("only 1D, 2D, 3D shifts supported")
    return n, t, c, h, w


def _check_dtypes(x, shift):
    if x.dtype not in _FLOAT_TYPES:
        raise ValueError("rubiks_shift_3d only supports float32, float64, float16 and bfloat16 dtypes.")
    low_precision = x.dtype in (torch.float16, torch.bfloat16)
    assert x.dtype == shift.dtype or (low_precision and shift.dtype == torch.float32), \
        "x and shift must have the same dtype (float32 shift is accepted with half / bfloat16 x)"


def rubiks_shift_3d_forward(x, shift, stride, padding, quantize=False, output=None):
    """Pure forward primitive, no autograd (primitive.py:54-80).  x: [N,T,C,H,W], shift: [3,C]."""
    strides, paddings = make_tuple(stride, 3), make_tuple(padding, 3)
    assert x.is_cuda, "rubiks shift only works on CUDA tensors"
    assert x.size(2) == shift.size(1), "x tensor channel dim[2] must match shift channel dim[1]"
    _check_dtypes(x, shift)
    output = allocate_output(output, x, compute_output_shape(x, strides, paddings, 3))
    ret = _native.rubiks_shift_3d_forward(x, shift, strides, paddings, quantize, output)
    assert ret == 0, "CUDA kernel return code {} != 0, error".format(ret)
    return output


def rubiks_shift_3d_backward(upstream_grad, x, shift, stride, padding, normalize_grad,
                             normalize_t_factor=1.0, quantize=False, x_grad_output=None,
                             shift_grad_output=None, need_x_grad=True, need_shift_grad=True):
    """Pure backward primitive (primitive.py:90-140): returns (x_grad, shift_grad).

    shift_grad is L2-normalised per channel when normalize_grad (T component scaled by
    normalize_t_factor first; a negative factor keeps only sign(grad_T))."""
    strides, paddings = make_tuple(stride, 3), make_tuple(padding, 3)
    assert x.is_cuda and upstream_grad.is_cuda, "rubiks shift only works on CUDA tensors"
    _check_dtypes(x, shift)
    x_grad = allocate_output(x_grad_output, x, x.size()) if need_x_grad else None
    shift_grad = allocate_output(shift_grad_output, shift, shift.size()) if need_shift_grad else None
    ret = _native.rubiks_shift_3d_backward(x, shift, upstream_grad, strides, paddings, x_grad, shift_grad,
                                           normalize_grad, normalize_t_factor, quantize)
    assert ret == 0, "CUDA return code {} != 0, error".format(ret)
    return x_grad, shift_grad


class RubiksShift3DFunc(torch.autograd.Function):
    """autograd.Function of the 3D shift (primitive.py:148-190): 7 inputs
    (x, shift, stride, padding, normalize_grad, normalize_t_factor, quantize)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, shift, stride, padding, normalize_grad, normalize_t_factor, quantize):
        assert isinstance(normalize_grad, bool)
        ctx.stride, ctx.padding = stride, padding
        ctx.normalize_grad, ctx.normalize_t_factor, ctx.quantize = normalize_grad, normalize_t_factor, quantize
        x = x.contiguous()
        shift = shift.contiguous()
        ctx.save_for_backward(x, shift)
        return rubiks_shift_3d_forward(x, shift, stride, padding, quantize=quantize)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        x, shift = ctx.saved_tensors
        x_grad = shift_grad = None
        need_x, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_x or need_s:
            x_grad, shift_grad = rubiks_shift_3d_backward(
                grad_output.contiguous(), x, shift, stride=ctx.stride, padding=ctx.padding,
                normalize_grad=ctx.normalize_grad, normalize_t_factor=ctx.normalize_t_factor,
                quantize=ctx.quantize, need_x_grad=need_x, need_shift_grad=need_s)
        return x_grad, shift_grad, None, None, None, None, None


# BASELINE.json / north_star call the op "RubiksShift3DFunction"
RubiksShift3DFunction = RubiksShift3DFunc


def rubiks_shift_3d(x, shift, stride=1, padding=0, normalize_grad=True, normalize_t_factor=1.0,
                    quantize=False):
    """User-facing differentiable 3D shift (primitive.py:193-215)."""
    assert len(x.size()) == 5, "x must be [N, T, C, H, W]"
    _, frames, channels, height, _ = x.size()
    assert channels == shift.size(1), "group shift is deprecated. Now C dim must match."
    if normalize_t_factor == "auto":
        normalize_t_factor = frames / height  # primitive.py:209-210
    else:
        assert isinstance(normalize_t_factor, (int, float))
    return RubiksShift3DFunc.apply(x, shift, stride, padding, normalize_grad, normalize_t_factor, quantize)
