"""2D learnable shift over [N,C,H,W]: primitives, autograd function (VFS2DFunc), functional.

Host-side mirror of rubiksnet/shiftlib/rubiks2d/primitive.py over librubiks_b200's C ABI.
"""
import torch

from ... import rubiksnet_cuda as _native
from ...utils import allocate_output, make_tuple

__all__ = ["rubiks2d", "rubiks2d_forward", "rubiks2d_backward", "VFS2DFunc", "compute_output_shape"]


def compute_output_shape(x, stride, padding, shift_dim=2):
    """primitive.py:13-23."""
    assert shift_dim == 2, "TODO"
    n, c, h, w = x.size()
    s, p = make_tuple(stride, 2), make_tuple(padding, 2)
    return n, c, (h + 2 * p[0] - 1) // s[0] + 1, (w + 2 * p[1] - 1) // s[1] + 1


def _check_dtypes(x, shift):
    low_precision = x.dtype in (torch.float16, torch.bfloat16)
    assert x.dtype == shift.dtype or (low_precision and shift.dtype == torch.float32), \
        "x and shift must have the same dtype (float32 shift is accepted with half / bfloat16 x)"


def rubiks2d_forward(x, shift, stride=1, padding=0, quantize=False, output=None):
    """Pure forward primitive (primitive.py:44-66).  x: [N,C,H,W], shift: [2,C] rows (H,W)."""
    strides, paddings = make_tuple(stride, 2), make_tuple(padding, 2)
    assert x.is_cuda, "shift only works on CUDA tensors"
    _check_dtypes(x, shift)
    output = allocate_output(output, x, compute_output_shape(x, strides, paddings))
    ret = _native.rubiks2d_forward(input=x, shift=shift, strides=strides, paddings=paddings,
                                   quantize=quantize, output=output)
    assert ret == 0, "CUDA kernel return code {} != 0, error".format(ret)
    return output


def rubiks2d_backward(upstream_grad, x, shift, stride, padding, normalize_grad=True, enable_shift_grad=True,
                      quantize=False, x_grad_output=None, shift_grad_output=None):
    """Pure backward primitive (primitive.py:77-122): returns (x_grad, shift_grad).  With
    enable_shift_grad=False shift_grad is all zeros (the reference leaves its new_zeros buffer untouched)."""
    strides, paddings = make_tuple(stride, 2), make_tuple(padding, 2)
    assert x.is_cuda and upstream_grad.is_cuda and shift.is_cuda, "shift only works on CUDA tensors"
    _check_dtypes(x, shift)
    x_grad = allocate_output(x_grad_output, x, x.size())
    if shift_grad_output is None and not enable_shift_grad:
        shift_grad = torch.zeros_like(shift)
    else:
        shift_grad = allocate_output(shift_grad_output, shift, shift.size())
    ret = _native.rubiks2d_backward(upstream_grad=upstream_grad, input=x, shift=shift, strides=strides,
                                    paddings=paddings, normalize_grad=normalize_grad,
                                    enable_shift_grad=enable_shift_grad, quantize=quantize,
                                    input_grad=x_grad, shift_grad=shift_grad)
    assert ret == 0, "CUDA return code {} != 0, error".format(ret)
    return x_grad, shift_grad


class VFS2DFunc(torch.autograd.Function):
    """autograd.Function of the 2D shift (primitive.py:131-174): inputs
    (x, shift, stride, padding, normalize_grad, enable_shift_grad, quantize)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, shift, stride, padding, normalize_grad, enable_shift_grad, quantize):
        assert isinstance(normalize_grad, bool)
        assert isinstance(enable_shift_grad, bool)
        ctx.stride, ctx.padding = stride, padding
        ctx.normalize_grad, ctx.enable_shift_grad, ctx.quantize = normalize_grad, enable_shift_grad, quantize
        x = x.contiguous()
        shift = shift.contiguous()
        ctx.save_for_backward(x, shift)
        return rubiks2d_forward(x, shift, stride, padding, quantize)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_output):
        x, shift = ctx.saved_tensors
        x_grad = shift_grad = None
        if any(ctx.needs_input_grad):
            gx, gs = rubiks2d_backward(grad_output.contiguous(), x, shift, stride=ctx.stride,
                                       padding=ctx.padding, normalize_grad=ctx.normalize_grad,
                                       enable_shift_grad=ctx.enable_shift_grad, quantize=ctx.quantize)
            if ctx.needs_input_grad[0]:
                x_grad = gx
            if ctx.needs_input_grad[1]:
                shift_grad = gs
        return x_grad, shift_grad, None, None, None, None, None


def rubiks2d(x, shift, stride=1, padding=0, normalize_grad=True, enable_shift_grad=True, quantize=False):
    """User-facing differentiable 2D shift (primitive.py:177-196)."""
    assert len(x.size()) == 4, "x must be [N, C, H, W]"
    return VFS2DFunc.apply(x, shift, stride, padding, normalize_grad, enable_shift_grad, quantize)
