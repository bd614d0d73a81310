"""Fused building blocks of RubiksShiftBlock (rubiksnet/backbone.py:74-135) on CUDA:

  bn_act(x, bn, relu)            BatchNorm2d (+ReLU) in two streaming passes of librubiks_b200 (bn_relu.cu)
  conv1x1(x, weight, residual)   the block's 1x1 convolutions as NCHW batched GEMMs  out[n] = W @ x[n] (+ residual[n])
                                 on cuBLAS (plain library GEMM; the residual add `out += shortcut` rides in the GEMM
                                 epilogue), avoiding cuDNN's NCHW<->NHWC transposes around every 1x1 conv

Both are autograd Functions over [N*T, C, H, W] activations; parameters stay fp32 (activations may be bf16).
"""
import torch

from . import _lib
from .rubiksnet_cuda import _on_device

__all__ = ["bn_act", "conv1x1"]


class _BNAct(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, relu):
        x = x.contiguous()
        ni, c = x.shape[0], x.shape[1]
        hw = x.numel() // max(ni * c, 1)
        y = torch.empty_like(x)
        mean_invstd = torch.empty(c, 2, dtype=torch.float32, device=x.device)
        scale_bias = torch.empty(c, 2, dtype=torch.float32, device=x.device)
        with _on_device(x.device):
            L = _lib.lib()
            nbytes = L.rb_bn_workspace_bytes(ni, c)
            ws = _lib.workspace(nbytes, x.device)
            _lib.check(L.rb_bn_act_forward(
                _lib.ptr(x), _lib.ptr(weight), _lib.ptr(bias), _lib.ptr(running_mean), _lib.ptr(running_var),
                _lib.ptr(y), _lib.ptr(mean_invstd), _lib.ptr(scale_bias), _lib.dtype_code(x), ni, c, hw,
                int(training), float(momentum), float(eps), int(relu), _lib.ptr(ws), nbytes,
                _lib.stream_handle(x.device)))
        ctx.save_for_backward(x, weight, mean_invstd, scale_bias)
        ctx.cfg = (ni, c, hw, bool(training), bool(relu))
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, weight, mean_invstd, scale_bias = ctx.saved_tensors
        ni, c, hw, training, relu = ctx.cfg
        dy = dy.contiguous()
        need_x = ctx.needs_input_grad[0]
        need_w = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dx = torch.empty_like(x) if need_x else None
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device) if need_w else None
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device) if need_w else None
        with _on_device(x.device):
            L = _lib.lib()
            nbytes = L.rb_bn_workspace_bytes(ni, c)
            ws = _lib.workspace(nbytes, x.device)
            _lib.check(L.rb_bn_act_backward(
                _lib.ptr(x), _lib.ptr(dy), None, _lib.ptr(weight), _lib.ptr(mean_invstd), _lib.ptr(scale_bias),
                _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.dtype_code(x), ni, c, hw, int(training),
                int(relu), _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
        return dx, dgamma, dbeta, None, None, None, None, None, None


def bn_act(x, bn, relu=True):
    """relu(bn(x)) with torch.nn.BatchNorm2d semantics (batch statistics + running-stat update in training
    mode, running statistics in eval mode) for a CUDA NCHW tensor."""
    training = bn.training or bn.running_mean is None
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = 0.1 if bn.momentum is None else bn.momentum
    w = bn.weight.float() if bn.weight is not None and bn.weight.dtype != torch.float32 else bn.weight
    b = bn.bias.float() if bn.bias is not None and bn.bias.dtype != torch.float32 else bn.bias
    return _BNAct.apply(x, w, b, bn.running_mean if bn.track_running_stats else None,
                        bn.running_var if bn.track_running_stats else None, training, momentum, bn.eps, relu)


class _Conv1x1(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, residual):
        ni, cin = x.shape[0], x.shape[1]
        cout = weight.shape[0]
        hshape = x.shape[2:]
        xb = x.contiguous().view(ni, cin, -1)
        wb = weight.view(cout, cin).to(x.dtype)
        if residual is None:
            out = torch.matmul(wb, xb)
        else:
            out = torch.baddbmm(residual.contiguous().view(ni, cout, -1), wb.unsqueeze(0).expand(ni, cout, cin), xb)
        ctx.save_for_backward(xb, wb)
        ctx.has_res = residual is not None
        ctx.wshape = weight.shape
        ctx.wdtype = weight.dtype
        return out.view(ni, cout, *hshape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        xb, wb = ctx.saved_tensors
        ni, cin, hw = xb.shape
        cout = wb.shape[0]
        gb = g.contiguous().view(ni, cout, hw)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = torch.matmul(wb.t(), gb).view(ni, cin, *g.shape[2:])
        if ctx.needs_input_grad[1]:
            # per-image [Cout,Cin] products summed over the batch in fp32
            gw = torch.bmm(gb, xb.transpose(1, 2)).sum(0, dtype=torch.float32).view(ctx.wshape).to(ctx.wdtype)
        gres = g if ctx.has_res and ctx.needs_input_grad[2] else None
        return gx, gw, gres


def conv1x1(x, weight, residual=None, stride=1):
    """nn.Conv2d(k=1, bias=False) on NCHW `x` with weight [Cout,Cin,1,1]; `residual` is added in the GEMM
    epilogue; stride > 1 sub-samples x first (backbone.py:104-105 shortcut)."""
    if stride != 1:
        x = x[:, :, ::stride, ::stride]
    return _Conv1x1.apply(x, weight, residual)
