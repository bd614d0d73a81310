"""AttentionShift (attention-quantized temporal shift), mirror of rubiksnet/attention_shift.py.

The reference expresses the 3-tap temporal mix as a depth-wise F.conv1d with C*H*W groups on a
transposed view (several HBM passes, :32-39).  Here the [C,3] tap softmax stays in PyTorch (it is 3*C
numbers and carries the gradient to `weight`), and the mix itself is one librubiks_b200 kernel that
reads x once and writes the output once.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .rubiksnet_cuda import _on_device

__all__ = ["AttentionShift", "attention_shift_mix"]


def attention_mix_forward(x, taps, n_segment, in_scale_bias=None):
    """Plain (no autograd) forward of the 3-tap temporal mix: x [N*T, C, H, W] contiguous, taps fp32 [C, 3].
    in_scale_bias fp32 [C, 2]: mix relu(x * scale + bias) instead of x (bn1 -> relu folded into the load)."""
    assert x.is_cuda, "attention shift only works on CUDA tensors"
    nt, c, h, w = x.shape
    assert nt % n_segment == 0, "batch (N*T) must be a multiple of n_segment"
    out = torch.empty_like(x)
    with _on_device(x.device), _lib.timed("attention_shift_forward", _lib.nbytes(x, out)):
        if in_scale_bias is None:
            _lib.check(_lib.lib().rb_attention_shift_forward(
                _lib.ptr(x), _lib.ptr(taps), _lib.ptr(out), _lib.dtype_code(x), nt // n_segment, n_segment,
                c, h * w, _lib.stream_handle(x.device)))
        else:
            _lib.check(_lib.lib().rb_bn_attention_shift_forward(
                _lib.ptr(x), _lib.ptr(in_scale_bias), _lib.ptr(taps), _lib.ptr(out), _lib.dtype_code(x), nt // n_segment,
                n_segment, c, h * w, _lib.stream_handle(x.device)))
    return out


def attention_mix_backward(x, taps, grad_out, n_segment, need_x=True, need_t=True, in_scale_bias=None):
    """(x_grad, taps_grad) of attention_mix_forward; either may be skipped.  With in_scale_bias, x is the tensor in front of
    the folded bn -> relu and x_grad the gradient with respect to the normalised input."""
    nt, c, h, w = x.shape
    n = nt // n_segment
    gx = torch.empty_like(x) if need_x else None
    gt = torch.empty_like(taps) if need_t else None
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_attention_shift_backward_workspace_bytes(n, n_segment, c, h * w) if need_t else 0
        ws = _lib.workspace(nbytes, x.device)
        with _lib.timed("attention_shift_backward", _lib.nbytes(x, grad_out, gx)):
            if in_scale_bias is None:
                _lib.check(L.rb_attention_shift_backward(
                    _lib.ptr(x), _lib.ptr(taps), _lib.ptr(grad_out), _lib.ptr(gx), _lib.ptr(gt), _lib.dtype_code(x),
                    n, n_segment, c, h * w, _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
            else:
                _lib.check(L.rb_bn_attention_shift_backward(
                    _lib.ptr(x), _lib.ptr(in_scale_bias), _lib.ptr(taps), _lib.ptr(grad_out), _lib.ptr(gx), _lib.ptr(gt),
                    _lib.dtype_code(x), n, n_segment, c, h * w, _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
    return gx, gt


class _AttentionMix(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, taps, n_segment):
        x = x.contiguous()
        taps = taps.contiguous().float()
        out = attention_mix_forward(x, taps, n_segment)
        ctx.save_for_backward(x, taps)
        ctx.n_segment = n_segment
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        x, taps = ctx.saved_tensors
        need_x, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_x or need_t):
            return None, None, None
        gx, gt = attention_mix_backward(x, taps, grad_out.contiguous(), ctx.n_segment, need_x, need_t)
        return gx, gt, None


def attention_shift_mix(x, taps, n_segment):
    """out[n,t,c] = taps[c,0] x[n,t-1,c] + taps[c,1] x[n,t,c] + taps[c,2] x[n,t+1,c]; x is [N*T,C,H,W]."""
    return _AttentionMix.apply(x, taps, n_segment)


class AttentionShift(nn.Module):
    """Per-channel 3-tap temporal attention.  State-dict keys: ``T`` (temperature, not trained) and
    ``weight`` [C,3].  Like the reference, ``weight`` is created on the first forward when `channels`
    is not given (attention_shift.py:24-27); passing `channels` creates it eagerly so that a checkpoint
    can be loaded without a dummy CUDA forward."""

    def __init__(self, n_segment, channels=None):
        super().__init__()
        self.n_segment = n_segment
        self.kernel_size = 3
        self.T = nn.Parameter(torch.tensor(2.0), requires_grad=False)
        self.weight = None if channels is None else nn.Parameter(torch.rand(channels, self.kernel_size))

    def taps(self):
        # attention_shift.py:29-30
        w = self.weight / (torch.std(self.weight, dim=1, keepdim=True) + 1e-6)
        return F.softmax(w / self.T, dim=1)

    def forward(self, x):
        if self.weight is None:
            self.weight = nn.Parameter(torch.rand(x.size(1), self.kernel_size).to(x.device))
        return attention_shift_mix(x, self.taps().float(), self.n_segment)

    attention_shift = forward
