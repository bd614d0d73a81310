"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the RubiksNet shift hot path.

A plain-C restatement (oracle/rubiks_oracle_impl.h, every function cites the reference lines it
follows) of ``cuda_src/rubiks3d_kernels.cu`` / ``cuda_src/rubiks2d_kernels.cu`` plus a numpy
restatement of ``rubiksnet/attention_shift.py``.  The reference ships no CPU path at all
(``rubiksnet/shiftlib/rubiks3d/primitive.py:61`` asserts ``x.is_cuda``), so this oracle is pinned
against outputs of the reference's own CUDA extension recorded on a B200 (tests/golden/, generated
by tests/golden/make_golden.py).

Import rules: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``rubiksnet_b200`` never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librubiks_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        assert _lib.oracle_abi_version() == 1
    return _lib


def set_num_threads(n):
    """Number of OpenMP threads the oracle uses (for the cpu_baseline timing)."""
    lib()
    ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))


def _triple(v, n):
    if isinstance(v, int):
        return [v] * n
    v = [int(e) for e in v]
    assert len(v) == n
    return v


def out_len(size, stride, pad):
    """cuda_src/rubiks.cpp:161-178 compute_output_shape (C++ int division, operands >= 0)."""
    return (size + 2 * pad - 1) // stride + 1


def _np(a, dtype=None):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a = np.ascontiguousarray(a)
    if dtype is not None:
        a = a.astype(dtype, copy=False)
    return a


def _sfx(dtype):
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise ValueError("oracle supports float32 / float64 only, got %s" % dtype)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------------------- 3D


def shift3d_forward(x, shift, stride=1, padding=0, quantize=False):
    x = _np(x)
    shift = _np(shift, x.dtype)
    sfx, _ = _sfx(x.dtype)
    N, T, C, H, W = x.shape
    assert shift.shape == (3, C)
    s, p = _triple(stride, 3), _triple(padding, 3)
    To, Ho, Wo = out_len(T, s[0], p[0]), out_len(H, s[1], p[1]), out_len(W, s[2], p[2])
    out = np.zeros((N, To, C, Ho, Wo), dtype=x.dtype)
    getattr(lib(), "oracle_shift3d_forward" + sfx)(
        _p(x), _p(shift), _p(out), N, T, C, H, W, To, Ho, Wo, *s, *p, int(bool(quantize)))
    return out


def shift3d_backward(x, shift, og, stride=1, padding=0, normalize_grad=True,
                     normalize_t_factor=1.0, quantize=False):
    """Returns (x_grad, shift_grad) like rubiks_shift_3d_backward (primitive.py:90-140)."""
    x = _np(x)
    shift = _np(shift, x.dtype)
    og = _np(og, x.dtype)
    sfx, cty = _sfx(x.dtype)
    N, T, C, H, W = x.shape
    s, p = _triple(stride, 3), _triple(padding, 3)
    To, Ho, Wo = out_len(T, s[0], p[0]), out_len(H, s[1], p[1]), out_len(W, s[2], p[2])
    assert og.shape == (N, To, C, Ho, Wo), (og.shape, (N, To, C, Ho, Wo))
    gin = np.zeros_like(x)
    gshift = np.zeros_like(shift)
    getattr(lib(), "oracle_shift3d_backward" + sfx)(
        _p(x), _p(shift), _p(og), _p(gin), _p(gshift), N, T, C, H, W, To, Ho, Wo, *s, *p,
        int(bool(normalize_grad)), cty(float(normalize_t_factor)), int(bool(quantize)))
    return gin, gshift


def shift3d_backward_shift_raw(x, shift, og, stride=1, padding=0):
    """Un-normalised shift gradient [3,C] only."""
    x = _np(x)
    shift = _np(shift, x.dtype)
    og = _np(og, x.dtype)
    sfx, _ = _sfx(x.dtype)
    N, T, C, H, W = x.shape
    s, p = _triple(stride, 3), _triple(padding, 3)
    To, Ho, Wo = og.shape[1], og.shape[3], og.shape[4]
    gshift = np.zeros_like(shift)
    getattr(lib(), "oracle_shift3d_backward_shift" + sfx)(
        _p(x), _p(shift), _p(og), _p(gshift), N, T, C, H, W, To, Ho, Wo, *s, *p)
    return gshift


def normalize_shift_grad_3d(g, factor=1.0):
    g = _np(g).copy()
    sfx, cty = _sfx(g.dtype)
    getattr(lib(), "oracle_normalize_shift_grad_3d" + sfx)(_p(g), g.shape[1], cty(float(factor)))
    return g


# ----------------------------------------------------------------------------------------- 2D


def shift2d_forward(x, shift, stride=1, padding=0, quantize=False):
    x = _np(x)
    shift = _np(shift, x.dtype)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    assert shift.shape == (2, C)
    s, p = _triple(stride, 2), _triple(padding, 2)
    Ho, Wo = out_len(H, s[0], p[0]), out_len(W, s[1], p[1])
    out = np.zeros((N, C, Ho, Wo), dtype=x.dtype)
    getattr(lib(), "oracle_shift2d_forward" + sfx)(
        _p(x), _p(shift), _p(out), N, C, H, W, Ho, Wo, *s, *p, int(bool(quantize)))
    return out


def shift2d_backward(x, shift, og, stride=1, padding=0, normalize_grad=True,
                     enable_shift_grad=True, quantize=False):
    x = _np(x)
    shift = _np(shift, x.dtype)
    og = _np(og, x.dtype)
    sfx, _ = _sfx(x.dtype)
    N, C, H, W = x.shape
    s, p = _triple(stride, 2), _triple(padding, 2)
    Ho, Wo = og.shape[2], og.shape[3]
    gin = np.zeros_like(x)
    gshift = np.zeros_like(shift)
    getattr(lib(), "oracle_shift2d_backward" + sfx)(
        _p(x), _p(shift), _p(og), _p(gin), _p(gshift), N, C, H, W, Ho, Wo, *s, *p,
        int(bool(normalize_grad)), int(bool(enable_shift_grad)), int(bool(quantize)))
    return gin, gshift


# --------------------------------------------------------------------------- AttentionShift


def attention_taps(weight, temperature):
    """rubiksnet/attention_shift.py:29-30: softmax((W / (std(W, dim=1, unbiased) + 1e-6)) / T)."""
    w = _np(weight).astype(np.float64)
    std = w.std(axis=1, ddof=1, keepdims=True)
    z = (w / (std + 1e-6)) / float(temperature)
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=1, keepdims=True)


def attention_shift_forward(x, weight, temperature, n_segment):
    """rubiksnet/attention_shift.py:18-39: out[n,t,c] = a[c,0] x[t-1] + a[c,1] x[t] + a[c,2] x[t+1]
    (zero beyond the clip); x is [N*T, C, H, W]."""
    x = _np(x)
    a = attention_taps(weight, temperature).astype(x.dtype)
    nt, C, H, W = x.shape
    xv = x.reshape(nt // n_segment, n_segment, C, H, W)
    out = a[None, None, :, 1, None, None] * xv
    out[:, 1:] += a[None, None, :, 0, None, None] * xv[:, :-1]
    out[:, :-1] += a[None, None, :, 2, None, None] * xv[:, 1:]
    return out.reshape(nt, C, H, W)


def attention_shift_backward(x, weight, temperature, n_segment, og):
    """Adjoint of attention_shift_forward wrt x and wrt the raw weight [C,3] (T is not trained:
    rubiksnet/attention_shift.py:11).  float64 throughout."""
    x = _np(x).astype(np.float64)
    og = _np(og).astype(np.float64)
    w = _np(weight).astype(np.float64)
    a = attention_taps(w, temperature)
    nt, C, H, W = x.shape
    xv = x.reshape(nt // n_segment, n_segment, C, H, W)
    gv = og.reshape(xv.shape)
    gx = a[None, None, :, 1, None, None] * gv
    gx[:, :-1] += a[None, None, :, 0, None, None] * gv[:, 1:]
    gx[:, 1:] += a[None, None, :, 2, None, None] * gv[:, :-1]
    ga = np.zeros_like(a)
    ga[:, 1] = (gv * xv).sum(axis=(0, 1, 3, 4))
    ga[:, 0] = (gv[:, 1:] * xv[:, :-1]).sum(axis=(0, 1, 3, 4))
    ga[:, 2] = (gv[:, :-1] * xv[:, 1:]).sum(axis=(0, 1, 3, 4))
    # softmax backward: gz = a * (ga - sum(ga * a))
    gz = a * (ga - (ga * a).sum(axis=1, keepdims=True))
    # z = w / ((std + eps) * T);  std = sqrt(sum((w-m)^2)/(K-1))
    K = w.shape[1]
    m = w.mean(axis=1, keepdims=True)
    std = w.std(axis=1, ddof=1, keepdims=True)
    d = (std + 1e-6) * float(temperature)
    dstd_dw = (w - m) / ((K - 1) * np.maximum(std, 1e-300))
    gw = gz / d - (gz * w).sum(axis=1, keepdims=True) * float(temperature) / (d * d) * dstd_dw
    return gx.reshape(nt, C, H, W), gw
