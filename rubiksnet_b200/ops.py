"""Raw (non-autograd) calls into librubiks_b200's block-level kernels: BatchNorm statistics / apply / backward, the
tcgen05 pointwise-conv GEMMs (forward, input gradient, weight gradient) and the fused 3D-shift + conv3 launches.

Everything here takes contiguous CUDA tensors [NI, C, H, W] (NI = clips * frames), launches on torch's current
stream and returns torch tensors; there is no fallback.  When ``timing.enabled`` every call is bracketed by CUDA
events on the launching stream so that bench.py can attribute time and algorithmic bytes to kernels.
"""
import torch

from . import _lib
from .rubiksnet_cuda import _on_device

BF16 = torch.bfloat16


timing = _lib.timing
_timed = _lib.timed


def _nbytes(*tensors):
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


def _wdt(w):
    if w.dtype == torch.float32:
        return _lib.RB_F32
    if w.dtype == BF16:
        return _lib.RB_BF16
    raise ValueError("conv weight must be float32 or bfloat16, got %s" % w.dtype)


# ------------------------------------------------------------------------------------------ BatchNorm

def bn_forward(x, gamma, beta, running_mean, running_var, training, momentum, eps, relu=True, apply=True):
    """Statistics (+ apply pass when `apply`) of BatchNorm2d(+ReLU).  Returns (y or None, mean_invstd, scale_bias)."""
    ni, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(ni * c, 1)
    y = torch.empty_like(x) if apply else None
    mean_invstd = torch.empty(c, 2, dtype=torch.float32, device=x.device)
    scale_bias = torch.empty(c, 2, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_bn_workspace_bytes(ni, c)
        ws = _lib.workspace(nbytes, x.device)
        with _timed("bn_forward<%s>" % ("stats+apply" if apply else "stats"), _nbytes(x) * (2 if apply else 1) + _nbytes(y)):
            _lib.check(L.rb_bn_act_forward(
                _lib.ptr(x), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(y),
                _lib.ptr(mean_invstd), _lib.ptr(scale_bias), _lib.dtype_code(x), ni, c, hw, int(training), float(momentum),
                float(eps), int(relu), _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
    return y, mean_invstd, scale_bias


def bn_backward(x, dy, residual, gamma, mean_invstd, scale_bias, training, relu=True, need_dx=True, need_params=True):
    """dx (+ residual), dgamma, dbeta of BatchNorm2d(+ReLU); the ReLU mask is recomputed from x."""
    ni, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(ni * c, 1)
    dx = torch.empty_like(x) if need_dx else None
    dgamma = torch.empty(c, dtype=torch.float32, device=x.device) if need_params else None
    dbeta = torch.empty(c, dtype=torch.float32, device=x.device) if need_params else None
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_bn_workspace_bytes(ni, c)
        ws = _lib.workspace(nbytes, x.device)
        with _timed("bn_backward", 2 * _nbytes(x, dy) + _nbytes(residual, dx)):
            _lib.check(L.rb_bn_act_backward(
                _lib.ptr(x), _lib.ptr(dy), _lib.ptr(residual), _lib.ptr(gamma), _lib.ptr(mean_invstd), _lib.ptr(scale_bias),
                _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.dtype_code(x), ni, c, hw, int(training), int(relu),
                _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
    return dx, dgamma, dbeta


# ------------------------------------------------------------------------------------------ pointwise convs

def bn_finalize(stats, count, gamma, beta, running_mean, running_var, momentum, eps):
    """(mean_invstd, scale_bias) of training-mode BatchNorm2d from the partial sums a GEMM epilogue produced
    (`stats` = (partial fp64 [C, splits, 2], splits) as returned by pw_conv(..., stats=True)); updates the running
    statistics like nn.BatchNorm2d."""
    partial, splits = stats
    c = gamma.shape[0]
    mean_invstd = torch.empty(c, 2, dtype=torch.float32, device=partial.device)
    scale_bias = torch.empty(c, 2, dtype=torch.float32, device=partial.device)
    with _on_device(partial.device):
        with _timed("bn_finalize", 0):
            _lib.check(_lib.lib().rb_bn_stats_finalize(
                _lib.ptr(partial), int(splits), c, float(count), _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(running_mean),
                _lib.ptr(running_var), float(momentum), float(eps), _lib.ptr(mean_invstd), _lib.ptr(scale_bias),
                _lib.stream_handle(partial.device)))
    return mean_invstd, scale_bias


def bn_apply(x, scale_bias, relu=True):
    """y = relu(x * scale + bias) with given per-channel coefficients [C, 2] (one streaming pass)."""
    ni, c = x.shape[0], x.shape[1]
    hw = x.numel() // max(ni * c, 1)
    y = torch.empty_like(x)
    with _on_device(x.device):
        with _timed("bn_apply", _nbytes(x, y)):
            _lib.check(_lib.lib().rb_bn_apply_forward(_lib.ptr(x), _lib.ptr(scale_bias), _lib.ptr(y), _lib.dtype_code(x), ni, c, hw,
                                                      int(relu), _lib.stream_handle(x.device)))
    return y


_MAX_STAT_SPLITS = 2 * 148


class WeightImage:
    """Packed shared-memory image of a [rows x contraction] bf16 weight matrix (rb_pw_weight_image_pack): what the
    second-generation kernel (csrc/pw_conv2.cu) bulk-copies per output-channel slice."""

    __slots__ = ("image", "rows", "contraction")

    def __init__(self, image, rows, contraction):
        self.image, self.rows, self.contraction = image, rows, contraction


def pw_image_supported(ni, k, n, hw, has_bn=False, preferred=False):
    """True when rb_pw_conv_forward has an image path for [ni, k, hw] -> [ni, n, hw] (preferred=True: ... and it is the
    faster schedule for this geometry)."""
    level = _lib.lib().rb_pw_conv_image_supported(int(ni), int(k), int(n), int(hw), int(bool(has_bn)))
    return level >= (2 if preferred else 1)


def pw_weight_images(weight):
    """(forward image, input-gradient image) of a fp32 conv weight [N,K(,1,1)]: pw_conv(x, fwd) is the convolution,
    pw_conv(g, dgrad) its input gradient."""
    assert weight.dtype == torch.float32 and weight.is_contiguous()
    n, k = weight.shape[0], weight.shape[1]
    L = _lib.lib()
    fwd = torch.empty(L.rb_pw_weight_image_bytes(n, k), dtype=torch.uint8, device=weight.device)
    bwd = torch.empty(L.rb_pw_weight_image_bytes(k, n), dtype=torch.uint8, device=weight.device)
    with _on_device(weight.device):
        with _timed("pw_weight_pack", 8 * n * k + fwd.numel() + bwd.numel()):
            _lib.check(L.rb_pw_weight_image_pack(_lib.ptr(weight), n, k, 0, _lib.ptr(fwd), _lib.stream_handle(weight.device)))
            _lib.check(L.rb_pw_weight_image_pack(_lib.ptr(weight), n, k, 1, _lib.ptr(bwd), _lib.stream_handle(weight.device)))
    return WeightImage(fwd, n, k), WeightImage(bwd, k, n)


def pw_conv(x, weight, residual=None, in_scale_bias=None, transposed=False, name="pw_conv", stats=False, resident=False):
    """out[i,n,p] = sum_k W[n,k] A[i,k,p] (+ residual).  weight: Conv2d parameter [N,K,1,1] / [N,K] (fp32 or bf16), or a
    WeightImage (second-generation TMA kernel); transposed=True reads a plain weight as [K,N] (input gradient of that
    conv).  A = relu(x*scale+bias) with in_scale_bias.  resident=True: `weight` came out of pw_weight_pack (RB_W_RESIDENT: the
    kernel stages it while the previous kernel of the stream is still draining)."""
    assert x.dtype == BF16 and x.is_contiguous()
    rflag = _lib.RB_W_RESIDENT if resident else 0
    if isinstance(weight, WeightImage):
        assert not transposed and not stats
        ni, k, n = x.shape[0], x.shape[1], weight.rows
        assert weight.contraction == k, "channel mismatch"
        hw = x.numel() // max(ni * k, 1)
        out = torch.empty((ni, n) + tuple(x.shape[2:]), dtype=BF16, device=x.device)
        if residual is not None:
            assert residual.dtype == BF16 and residual.is_contiguous() and residual.shape == out.shape
        with _on_device(x.device):
            with _timed(name, _nbytes(x, residual, out), 2 * ni * hw * k * n):
                _lib.check(_lib.lib().rb_pw_conv_forward(
                    _lib.ptr(x), _lib.ptr(weight.image), _lib.RB_W_IMAGE, 0, _lib.ptr(residual), _lib.ptr(out), _lib.RB_BF16,
                    ni, k, n, hw, _lib.ptr(in_scale_bias), _lib.stream_handle(x.device)))
        return out
    ni, k = x.shape[0], x.shape[1]
    n = weight.shape[1] if transposed else weight.shape[0]
    assert (weight.shape[0] if transposed else weight.shape[1]) == k, "channel mismatch"
    hw = x.numel() // max(ni * k, 1)
    out = torch.empty((ni, n) + tuple(x.shape[2:]), dtype=BF16, device=x.device)
    if residual is not None:
        assert residual.dtype == BF16 and residual.is_contiguous() and residual.shape == out.shape
    with _on_device(x.device):
        with _timed(name, _nbytes(x, residual, out), 2 * ni * hw * k * n):
            if stats:
                # stats=True: also returns (partial, splits), the per-channel (sum, sum of squares) of `out` reduced in
                # the GEMM epilogue -- input of bn_finalize for the BatchNorm that consumes `out`
                import ctypes
                partial = torch.empty(n * _MAX_STAT_SPLITS * 2, dtype=torch.float64, device=x.device)
                splits = ctypes.c_int(0)
                _lib.check(_lib.lib().rb_pw_conv_forward_stats(
                    _lib.ptr(x), _lib.ptr(weight), _wdt(weight) | rflag, int(transposed), _lib.ptr(residual), _lib.ptr(out),
                    _lib.RB_BF16, ni, k, n, hw, _lib.ptr(in_scale_bias), _lib.ptr(partial), partial.numel() * 8,
                    ctypes.byref(splits), _lib.stream_handle(x.device)))
                return out, (partial, splits.value)
            _lib.check(_lib.lib().rb_pw_conv_forward(
                _lib.ptr(x), _lib.ptr(weight), _wdt(weight) | rflag, int(transposed), _lib.ptr(residual), _lib.ptr(out), _lib.RB_BF16,
                ni, k, n, hw, _lib.ptr(in_scale_bias), _lib.stream_handle(x.device)))
    return out


def pw_stats_preferred(ni, k, n, hw):
    """True when pw_conv(..., stats=True) reduces the BatchNorm statistics of its output at (almost) no cost for this geometry
    (the tensor-map schedule: 16-byte row pitch) -- prefer it to a separate statistics pass over the tensor."""
    return bool(_lib.lib().rb_pw_conv_stats_preferred(int(ni), int(k), int(n), int(hw)))


def pw_conv_f32(x, weight, residual=None, in_scale_bias=None, out_scale_bias=None, relu=False, resident=False,
                name="pw_conv_tf32"):
    """fp32 1x1 convolution on the tcgen05 kind::tf32 kernel (inference path):
    out = [relu](conv(A) * out_scale + out_bias) + residual,  A = x or relu(x * in_scale + in_bias).
    x [NI,K,H,W] fp32, weight [N,K(,1,1)] fp32; the coefficient tensors are fp32 [C,2] as returned by bn_forward."""
    assert x.dtype == torch.float32 and x.is_contiguous() and weight.dtype == torch.float32 and weight.is_contiguous()
    ni, k, n = x.shape[0], x.shape[1], weight.shape[0]
    assert weight.shape[1] == k, "channel mismatch"
    hw = x.numel() // max(ni * k, 1)
    out = torch.empty((ni, n) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == out.shape
    for sb, c in ((in_scale_bias, k), (out_scale_bias, n)):
        assert sb is None or (sb.dtype == torch.float32 and sb.is_contiguous() and sb.numel() == 2 * c)
    with _on_device(x.device):
        with _timed(name, _nbytes(x, residual, out), 2 * ni * hw * k * n):
            _lib.check(_lib.lib().rb_pw_conv_forward_f32(
                _lib.ptr(x), _lib.ptr(weight), _lib.ptr(residual), _lib.ptr(out), ni, k, n, hw, _lib.ptr(in_scale_bias),
                _lib.ptr(out_scale_bias), int(bool(relu)), _lib.RB_W_RESIDENT if resident else 0, _lib.stream_handle(x.device)))
    return out


def im2col3x3(x, stride, tpad, out_dtype):
    """Patch matrix [NI, tpad, Ho, Wo] of a 3x3 / padding-1 convolution of x [NI, Cin, H, W] (rows >= 9*Cin are zero)."""
    assert x.is_contiguous() and x.dim() == 4
    ni, cin, h, w = x.shape
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    cols = torch.empty(ni, tpad, ho, wo, dtype=out_dtype, device=x.device)
    with _on_device(x.device):
        with _timed("im2col3x3", _nbytes(x, cols)):
            _lib.check(_lib.lib().rb_im2col3x3(_lib.ptr(x), _lib.ptr(cols), _lib.dtype_code(x), _lib.dtype_code(cols), ni, cin, h, w,
                                               int(stride), int(tpad), _lib.stream_handle(x.device)))
    return cols


def pw_weight_pack(weight):
    """bf16 copies of a fp32 conv weight [N,K(,1,1)] in both orientations: (w_nk [N,K], w_kn [K,N]); one launch.
    pw_conv(x, w_nk) is the forward, pw_conv(g, w_kn) the input gradient (no transposing weight staging)."""
    assert weight.dtype == torch.float32 and weight.is_contiguous()
    n, k = weight.shape[0], weight.shape[1]
    packed = torch.empty(2, n * k, dtype=BF16, device=weight.device)
    with _on_device(weight.device):
        with _timed("pw_weight_pack", 4 * n * k + 4 * n * k):
            _lib.check(_lib.lib().rb_pw_weight_pack(_lib.ptr(weight), _lib.ptr(packed[0]), _lib.ptr(packed[1]), n, k,
                                                    _lib.stream_handle(weight.device)))
    return packed[0].view(n, k), packed[1].view(k, n)


def pw_conv_wgrad(out_grad, x, in_scale_bias=None, name="pw_conv_wgrad"):
    """fp32 [N,K] weight gradient of the conv whose forward was pw_conv(x, W, in_scale_bias=...)."""
    assert x.dtype == BF16 and out_grad.dtype == BF16 and x.is_contiguous() and out_grad.is_contiguous()
    ni, k, n = x.shape[0], x.shape[1], out_grad.shape[1]
    hw = x.numel() // max(ni * k, 1)
    dw = torch.empty(n, k, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_pw_conv_wgrad_workspace_bytes(ni, k, n, hw)
        ws = _lib.workspace(nbytes, x.device)
        with _timed(name, _nbytes(x, out_grad), 2 * ni * hw * k * n):
            _lib.check(L.rb_pw_conv_wgrad(_lib.ptr(out_grad), _lib.ptr(x), _lib.ptr(dw), _lib.RB_BF16, ni, k, n, hw,
                                          _lib.ptr(in_scale_bias), _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
    return dw


def shift3d_pw_conv(x, shift, weight, residual, frames):
    """conv3(RubiksShift3D(x)) + residual in one launch.  x [N*T, C, H, W] bf16 (viewed [N,T,C,H,W])."""
    assert x.dtype == BF16 and x.is_contiguous() and x.dim() == 4
    nt, c, h, w = x.shape
    assert nt % frames == 0
    cout = weight.shape[0]
    out = torch.empty(nt, cout, h, w, dtype=BF16, device=x.device)
    if residual is not None:
        assert residual.dtype == BF16 and residual.is_contiguous() and residual.shape == out.shape
    with _on_device(x.device):
        with _timed("shift3d_pw_conv", _nbytes(x, residual, out), 2 * nt * h * w * c * cout):
            _lib.check(_lib.lib().rb_shift3d_pw_conv_forward(
                _lib.ptr(x), _lib.ptr(shift), _lib.ptr(weight), _wdt(weight), _lib.ptr(residual), _lib.ptr(out), _lib.RB_BF16,
                _lib.dtype_code(shift), nt // frames, frames, c, h, w, cout, _lib.stream_handle(x.device)))
    return out


def shift3d_pw_conv_wgrad(out_grad, x, shift, frames):
    """fp32 [Cout, C] weight gradient of conv3 in shift3d_pw_conv (the shift is recomputed inside the kernel)."""
    assert x.dtype == BF16 and out_grad.dtype == BF16 and x.is_contiguous() and out_grad.is_contiguous()
    nt, c, h, w = x.shape
    cout = out_grad.shape[1]
    dw = torch.empty(cout, c, dtype=torch.float32, device=x.device)
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_pw_conv_wgrad_workspace_bytes(nt, c, cout, h * w)
        ws = _lib.workspace(nbytes, x.device)
        with _timed("shift3d_pw_conv_wgrad", _nbytes(x, out_grad), 2 * nt * h * w * c * cout):
            _lib.check(L.rb_shift3d_pw_conv_wgrad(
                _lib.ptr(out_grad), _lib.ptr(x), _lib.ptr(shift), _lib.ptr(dw), _lib.RB_BF16, _lib.dtype_code(shift),
                nt // frames, frames, c, h, w, cout, _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
    return dw


def shift3d_backward(x, shift, out_grad, frames, normalize_grad, normalize_t_factor, need_x=True, need_shift=True, stride=1):
    """x_grad / shift_grad of the 3D shift (temporal stride 1, spatial `stride`, no padding) over x [N*T, C, H, W]
    (rb_shift3d_backward)."""
    nt, c, h, w = x.shape
    n = nt // frames
    x_grad = torch.empty_like(x) if need_x else None
    shift_grad = torch.empty_like(shift) if need_shift else None
    dt = _lib.dtype_code(x)
    geo = (n, frames, c, h, w, 1, int(stride), int(stride), 0, 0, 0)
    with _on_device(x.device):
        L = _lib.lib()
        nbytes = L.rb_shift3d_backward_workspace_bytes(dt, *geo) if need_shift else 0
        ws = _lib.workspace(nbytes, x.device)
        with _timed("shift3d_backward", _nbytes(x, out_grad, x_grad)):
            _lib.check(L.rb_shift3d_backward(
                _lib.ptr(x), _lib.ptr(shift), _lib.ptr(out_grad), _lib.ptr(x_grad), _lib.ptr(shift_grad), dt,
                _lib.dtype_code(shift), *geo, int(bool(normalize_grad)), float(normalize_t_factor), 0, _lib.ptr(ws), nbytes,
                _lib.stream_handle(x.device)))
    return x_grad, shift_grad
