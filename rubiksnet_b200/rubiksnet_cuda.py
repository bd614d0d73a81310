"""Drop-in for the reference's native module ``rubiksnet_cuda`` (pybind11 exports at
/root/reference/cuda_src/rubiks.cpp:384-396): same six function names, argument order, keyword names
(2D) and return value (0), backed by librubiks_b200.so through the C ABI instead of ATen.

Additive entry points: ``*_half`` / ``*_bfloat16`` for the 3D shift (the reference is float/double
only, rubiksnet/shiftlib/rubiks3d/primitive.py:66-75) and dtype-generic ``rubiks_shift_3d_forward`` /
``rubiks_shift_3d_backward``.

Differences a caller can observe, all deliberate (see include/rubiks_b200.h):
  * launches go to torch's CURRENT stream under a device guard (the reference uses the legacy default
    stream and device 0's attributes);
  * every output element is written, so outputs need not be pre-zeroed;
  * ``shift`` may be float32 while activations are float16 / bfloat16.
"""
import torch

from . import _lib

__all__ = [
    "rubiks2d_forward", "rubiks2d_backward",
    "rubiks_shift_3d_forward_float", "rubiks_shift_3d_forward_double",
    "rubiks_shift_3d_backward_float", "rubiks_shift_3d_backward_double",
    "rubiks_shift_3d_forward_half", "rubiks_shift_3d_forward_bfloat16",
    "rubiks_shift_3d_backward_half", "rubiks_shift_3d_backward_bfloat16",
    "rubiks_shift_3d_forward", "rubiks_shift_3d_backward",
]


def _check_tensor(t, name):
    # TX_CHECK_TENSOR (cuda_src/utils.h:303-305): CUDA + contiguous, else a RuntimeError
    if not t.is_cuda:
        raise RuntimeError("TorchX AssertError: %s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("TorchX AssertError: %s must be contiguous" % name)


class _on_device:
    """Minimal device guard: switches only when the tensor lives on another device."""

    def __init__(self, device):
        self.idx = device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


# ------------------------------------------------------------------------------------------ 2D


def rubiks2d_forward(input, shift, strides, paddings, quantize, output):
    """rubiks.cpp:44-67."""
    _check_tensor(input, "input")
    _check_tensor(shift, "shift")
    _check_tensor(output, "output")
    N, C, H, W = input.shape
    if tuple(shift.shape) != (2, C):
        raise RuntimeError("ShapeException: rubiks shift expected shape = [2, %d]; actual shape = %s"
                           % (C, list(shift.shape)))
    with _on_device(input.device), _lib.timed("shift2d_forward", _lib.nbytes(input, output)):
        _lib.check(_lib.lib().rb_shift2d_forward(
            _lib.ptr(input), _lib.ptr(shift), _lib.ptr(output), _lib.dtype_code(input), _lib.dtype_code(shift),
            N, C, H, W, int(strides[0]), int(strides[1]), int(paddings[0]), int(paddings[1]),
            int(bool(quantize)), _lib.stream_handle(input.device)))
    return 0


def rubiks2d_backward(upstream_grad, input, shift, strides, paddings, normalize_grad, enable_shift_grad,
                      quantize, input_grad, shift_grad):
    """rubiks.cpp:94-155.  shift_grad is overwritten (untouched when enable_shift_grad is False)."""
    for t, name in ((upstream_grad, "output_grad"), (input, "input"), (shift, "shift_field"),
                    (input_grad, "input_grad"), (shift_grad, "shift_grad")):
        _check_tensor(t, name)
    N, C, H, W = input.shape
    sH, sW, pH, pW = int(strides[0]), int(strides[1]), int(paddings[0]), int(paddings[1])
    dt = _lib.dtype_code(input)
    with _on_device(input.device):
        L = _lib.lib()
        nbytes = L.rb_shift2d_backward_workspace_bytes(dt, N, C, H, W, sH, sW, pH, pW) if enable_shift_grad else 0
        ws = _lib.workspace(nbytes, input.device)
        with _lib.timed("shift2d_backward", _lib.nbytes(input, upstream_grad, input_grad)):
            _lib.check(L.rb_shift2d_backward(
                _lib.ptr(input), _lib.ptr(shift), _lib.ptr(upstream_grad), _lib.ptr(input_grad), _lib.ptr(shift_grad),
                dt, _lib.dtype_code(shift), N, C, H, W, sH, sW, pH, pW, int(bool(normalize_grad)),
                int(bool(enable_shift_grad)), int(bool(quantize)), _lib.ptr(ws), nbytes,
                _lib.stream_handle(input.device)))
    return 0


# ------------------------------------------------------------------------------------------ 3D


def _assert_contiguous(t, name):
    # AT_ASSERTM(x->is_contiguous(), ...) rubiks.cpp:212-222
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)


def rubiks_shift_3d_forward(input, shift, strides, paddings, quantize, output, _expect=None):
    """rubiks.cpp:181-253 for any supported dtype."""
    if _expect is not None and input.dtype != _expect:
        raise RuntimeError("expected scalar type %s but found %s" % (_expect, input.dtype))
    _assert_contiguous(input, "input_tensor_ptr")
    _assert_contiguous(shift, "shift_tensor_ptr")
    _assert_contiguous(output, "output_tensor_ptr")
    N, T, C, H, W = input.shape
    with _on_device(input.device), _lib.timed("shift3d_forward", _lib.nbytes(input, output)):
        _lib.check(_lib.lib().rb_shift3d_forward(
            _lib.ptr(input), _lib.ptr(shift), _lib.ptr(output), _lib.dtype_code(input), _lib.dtype_code(shift),
            N, T, C, H, W, int(strides[0]), int(strides[1]), int(strides[2]),
            int(paddings[0]), int(paddings[1]), int(paddings[2]), int(bool(quantize)),
            _lib.stream_handle(input.device)))
    return 0


def rubiks_shift_3d_backward(input, shift, output_grad, strides, paddings, input_grad, shift_grad,
                             normalize_grad, normalize_t_factor, quantize, _expect=None):
    """rubiks.cpp:256-379 for any supported dtype.  input_grad / shift_grad may be None to skip one."""
    if _expect is not None and input.dtype != _expect:
        raise RuntimeError("expected scalar type %s but found %s" % (_expect, input.dtype))
    for t, name in ((input, "input"), (shift, "shift"), (output_grad, "output_grad")):
        _assert_contiguous(t, name)
    N, T, C, H, W = input.shape
    geo = (N, T, C, H, W, int(strides[0]), int(strides[1]), int(strides[2]),
           int(paddings[0]), int(paddings[1]), int(paddings[2]))
    dt = _lib.dtype_code(input)
    with _on_device(input.device):
        L = _lib.lib()
        nbytes = L.rb_shift3d_backward_workspace_bytes(dt, *geo) if shift_grad is not None else 0
        ws = _lib.workspace(nbytes, input.device)
        with _lib.timed("shift3d_backward", _lib.nbytes(input, output_grad, input_grad)):
            _lib.check(L.rb_shift3d_backward(
                _lib.ptr(input), _lib.ptr(shift), _lib.ptr(output_grad), _lib.ptr(input_grad), _lib.ptr(shift_grad),
                dt, _lib.dtype_code(shift), *geo, int(bool(normalize_grad)), float(normalize_t_factor),
                int(bool(quantize)), _lib.ptr(ws), nbytes, _lib.stream_handle(input.device)))
    return 0


def _typed(fn, dtype, name):
    def wrapper(*args):
        return fn(*args, _expect=dtype)
    wrapper.__name__ = name
    wrapper.__doc__ = fn.__doc__
    return wrapper


rubiks_shift_3d_forward_float = _typed(rubiks_shift_3d_forward, torch.float32, "rubiks_shift_3d_forward_float")
rubiks_shift_3d_forward_double = _typed(rubiks_shift_3d_forward, torch.float64, "rubiks_shift_3d_forward_double")
rubiks_shift_3d_forward_half = _typed(rubiks_shift_3d_forward, torch.float16, "rubiks_shift_3d_forward_half")
rubiks_shift_3d_forward_bfloat16 = _typed(rubiks_shift_3d_forward, torch.bfloat16, "rubiks_shift_3d_forward_bfloat16")
rubiks_shift_3d_backward_float = _typed(rubiks_shift_3d_backward, torch.float32, "rubiks_shift_3d_backward_float")
rubiks_shift_3d_backward_double = _typed(rubiks_shift_3d_backward, torch.float64, "rubiks_shift_3d_backward_double")
rubiks_shift_3d_backward_half = _typed(rubiks_shift_3d_backward, torch.float16, "rubiks_shift_3d_backward_half")
rubiks_shift_3d_backward_bfloat16 = _typed(rubiks_shift_3d_backward, torch.bfloat16,
                                           "rubiks_shift_3d_backward_bfloat16")
