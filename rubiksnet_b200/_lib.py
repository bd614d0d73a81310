"""ctypes binding of librubiks_b200.so (the C ABI declared in include/rubiks_b200.h).

PyTorch is used here for device memory, streams and the device guard only: every entry point
receives raw device pointers (``tensor.data_ptr()``) and the current CUDA stream handle.
There is NO fallback: if the library is missing or a call fails, this raises.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librubiks_b200.so")

RB_F32, RB_F64, RB_F16, RB_BF16 = 0, 1, 2, 3
RB_W_IMAGE = 16  # weight argument of rb_pw_conv_forward is a packed image (rb_pw_weight_image_pack)
RB_W_RESIDENT = 256  # OR-ed into weight_dtype: the buffer came from rb_pw_weight_pack* (may be staged before the PDL wait)
RB_IMPL_AUTO, RB_IMPL_GENERIC, RB_IMPL_TILED, RB_IMPL_STRIP = 0, 1, 2, 3
_DTYPES = {torch.float32: RB_F32, torch.float64: RB_F64, torch.float16: RB_F16, torch.bfloat16: RB_BF16}

_lib = None


class RubiksCudaError(RuntimeError):
    """Raised for a non-zero rb_status_t (the reference raises RuntimeError from C++ exceptions)."""


def _declare(lib):
    vp, i, sz, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_double
    lib.rb_abi_version.restype = i
    lib.rb_last_error.restype = ctypes.c_char_p
    lib.rb_launch_count.restype = ctypes.c_uint64
    lib.rb_launch_count_reset.restype = None
    lib.rb_set_impl.argtypes = [i]
    lib.rb_set_impl.restype = None
    lib.rb_set_dependent_launch.argtypes = [i]
    lib.rb_set_dependent_launch.restype = None
    lib.rb_bn_set_resident.argtypes = [i]
    lib.rb_bn_set_resident.restype = None
    lib.rb_pw_conv_tma_set_enabled.argtypes = [i]
    lib.rb_pw_conv_tma_set_enabled.restype = None
    lib.rb_last_impl.restype = i
    lib.rb_out_len.argtypes = [i, i, i]
    lib.rb_out_len.restype = i
    lib.rb_shift3d_forward.argtypes = [vp, vp, vp, i, i] + [i] * 5 + [i] * 6 + [i, vp]
    lib.rb_shift3d_forward.restype = i
    lib.rb_shift3d_backward_workspace_bytes.argtypes = [i] + [i] * 5 + [i] * 6
    lib.rb_shift3d_backward_workspace_bytes.restype = sz
    lib.rb_shift3d_backward.argtypes = [vp] * 5 + [i, i] + [i] * 5 + [i] * 6 + [i, dbl, i, vp, sz, vp]
    lib.rb_shift3d_backward.restype = i
    lib.rb_shift2d_forward.argtypes = [vp, vp, vp, i, i] + [i] * 4 + [i] * 4 + [i, vp]
    lib.rb_shift2d_forward.restype = i
    lib.rb_shift2d_backward_workspace_bytes.argtypes = [i] + [i] * 4 + [i] * 4
    lib.rb_shift2d_backward_workspace_bytes.restype = sz
    lib.rb_shift2d_backward.argtypes = [vp] * 5 + [i, i] + [i] * 4 + [i] * 4 + [i, i, i, vp, sz, vp]
    lib.rb_shift2d_backward.restype = i
    lib.rb_attention_shift_forward.argtypes = [vp, vp, vp, i, i, i, i, i, vp]
    lib.rb_attention_shift_forward.restype = i
    lib.rb_attention_shift_backward_workspace_bytes.argtypes = [i, i, i, i]
    lib.rb_attention_shift_backward_workspace_bytes.restype = sz
    lib.rb_attention_shift_backward.argtypes = [vp] * 5 + [i, i, i, i, i, vp, sz, vp]
    lib.rb_attention_shift_backward.restype = i
    lib.rb_bn_attention_shift_forward.argtypes = [vp, vp, vp, vp, i, i, i, i, i, vp]
    lib.rb_bn_attention_shift_forward.restype = i
    lib.rb_bn_attention_shift_backward.argtypes = [vp] * 6 + [i, i, i, i, i, vp, sz, vp]
    lib.rb_bn_attention_shift_backward.restype = i
    fl = ctypes.c_float
    lib.rb_bn_workspace_bytes.argtypes = [i, i]
    lib.rb_bn_workspace_bytes.restype = sz
    lib.rb_bn_act_forward.argtypes = [vp] * 8 + [i, i, i, i, i, fl, fl, i, vp, sz, vp]
    lib.rb_bn_act_forward.restype = i
    lib.rb_bn_act_backward.argtypes = [vp] * 9 + [i, i, i, i, i, i, vp, sz, vp]
    lib.rb_bn_act_backward.restype = i
    lib.rb_plane_reduce.argtypes = [vp, vp, vp, i, i, i, fl, vp]
    lib.rb_plane_reduce.restype = i
    lib.rb_plane_scale.argtypes = [vp, vp, vp, vp, i, i, i, vp]
    lib.rb_plane_scale.restype = i
    lib.rb_pw_conv_forward.argtypes = [vp, vp, i, i, vp, vp, i, i, i, i, i, vp, vp]
    lib.rb_pw_conv_forward.restype = i
    lib.rb_pw_conv_forward_f32.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp, i, i, vp]
    lib.rb_pw_conv_forward_f32.restype = i
    lib.rb_im2col3x3.argtypes = [vp, vp, i, i, i, i, i, i, i, i, vp]
    lib.rb_im2col3x3.restype = i
    lib.rb_frames_to_clip.argtypes = [vp, vp, i, i, i, i, i, ctypes.POINTER(fl), ctypes.POINTER(fl), i, vp]
    lib.rb_frames_to_clip.restype = i
    lib.rb_pw_conv_forward_stats.argtypes = [vp, vp, i, i, vp, vp, i, i, i, i, i, vp, vp, sz, ctypes.POINTER(ctypes.c_int), vp]
    lib.rb_pw_conv_forward_stats.restype = i
    lib.rb_bn_stats_finalize.argtypes = [vp, i, i, dbl, vp, vp, vp, vp, fl, fl, vp, vp, vp]
    lib.rb_bn_stats_finalize.restype = i
    lib.rb_bn_apply_forward.argtypes = [vp, vp, vp, i, i, i, i, i, vp]
    lib.rb_bn_apply_forward.restype = i
    lib.rb_pw_weight_pack.argtypes = [vp, vp, vp, i, i, vp]
    lib.rb_pw_weight_pack.restype = i
    lib.rb_shift3d_pw_conv_forward.argtypes = [vp, vp, vp, i, vp, vp] + [i] * 8 + [vp]
    lib.rb_shift3d_pw_conv_forward.restype = i
    lib.rb_pw_conv_wgrad_workspace_bytes.argtypes = [i, i, i, i]
    lib.rb_pw_conv_wgrad_workspace_bytes.restype = sz
    lib.rb_pw_conv_wgrad.argtypes = [vp, vp, vp, i, i, i, i, i, vp, vp, sz, vp]
    lib.rb_pw_conv_wgrad.restype = i
    lib.rb_shift3d_pw_conv_wgrad.argtypes = [vp] * 4 + [i] * 8 + [vp, sz, vp]
    lib.rb_shift3d_pw_conv_wgrad.restype = i
    lib.rb_pw_weight_image_bytes.argtypes = [i, i]
    lib.rb_pw_weight_image_bytes.restype = sz
    lib.rb_pw_weight_image_pack.argtypes = [vp, i, i, i, vp, vp]
    lib.rb_pw_weight_image_pack.restype = i
    lib.rb_pw_weight_pack_multi.argtypes = [vp, i, vp]
    lib.rb_pw_weight_pack_multi.restype = i
    lib.rb_pw_weight_image_pack_multi.argtypes = [vp, i, vp]
    lib.rb_pw_weight_image_pack_multi.restype = i
    lib.rb_pw_conv_image_supported.argtypes = [i, i, i, i, i]
    lib.rb_pw_conv_image_supported.restype = i
    lib.rb_pw_conv_image_set_tuning.argtypes = [i, i, i]
    lib.rb_pw_conv_image_set_tuning.restype = None
    lib.rb_pw_conv_stats_preferred.argtypes = [i, i, i, i]
    lib.rb_pw_conv_stats_preferred.restype = i
    lib.rb_pw_conv_set_tuning.argtypes = [i]
    lib.rb_pw_conv_set_tuning.restype = None
    lib.rb_pw_conv_wgrad_set_tuning.argtypes = [i, i, i]
    lib.rb_pw_conv_wgrad_set_tuning.restype = None


def lib():
    """The loaded library.  Raises ImportError when it has not been built (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "librubiks_b200.so is not built: run `python -m rubiksnet_b200.build` "
                "(or __graft_entry__.build()); there is no CPU / PyTorch fallback for the shift kernels")
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        if handle.rb_abi_version() != 1:
            raise ImportError("librubiks_b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RubiksCudaError("librubiks_b200 status %d: %s" % (rc, lib().rb_last_error().decode()))


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise ValueError("rubiks shift supports float32/float64/float16/bfloat16, got %s" % t.dtype)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_handle(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


# one growing scratch buffer per (device, stream): stream-ordered reuse is safe because every kernel
# that touches it is enqueued on that same stream
_workspaces = {}
_workspaces_lock = threading.Lock()  # autograd runs backward on its own threads (and nn.DataParallel one per device)


def workspace(nbytes, device):
    if nbytes == 0:
        return None
    stream = torch.cuda.current_stream(device)
    if torch.cuda.is_current_stream_capturing():
        # a buffer that lives in a CUDA graph's private pool must not be handed to eager work later: allocate per call
        # inside a capture (the pool recycles it between replays) and keep it out of the cache
        return torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
    key = (device.index, stream.cuda_stream)
    with _workspaces_lock:
        buf = _workspaces.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            _workspaces[key] = buf
    return buf


class _Timing:
    """Per-kernel CUDA-event timing, switched on by bench.py for one extra step."""

    def __init__(self):
        self.enabled = False
        self.records = []  # (name, algorithmic bytes, flops, start event, end event)

    def start(self):
        self.enabled, self.records = True, []

    def stop(self):
        self.enabled = False
        torch.cuda.synchronize()
        agg = {}
        for name, nbytes, flops, e0, e1 in self.records:
            d = agg.setdefault(name, {"bytes": 0, "flops": 0, "ms": 0.0, "launches": 0})
            d["bytes"] += nbytes
            d["flops"] += flops
            d["ms"] += e0.elapsed_time(e1)
            d["launches"] += 1
        self.records = []
        return agg


timing = _Timing()


class timed:
    def __init__(self, name, nbytes, flops=0):
        self.name, self.nbytes, self.flops = name, nbytes, flops

    def __enter__(self):
        if timing.enabled:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if timing.enabled:
            self.e1.record()
            timing.records.append((self.name, self.nbytes, self.flops, self.e0, self.e1))
        return False



def nbytes(*tensors):
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


def launch_count():
    return int(lib().rb_launch_count())


def reset_launch_count():
    lib().rb_launch_count_reset()


def set_impl(impl):
    lib().rb_set_impl(int(impl))


def set_dependent_launch(enabled):
    """Programmatic dependent launch of every library kernel on (default) / off -- for A/B measurements."""
    lib().rb_set_dependent_launch(int(bool(enabled)))


def set_pw_tma(enabled):
    """Tensor-map TMA schedule of the plain 1x1 convolutions on large maps on (default) / off -- A/B measurements, tests."""
    lib().rb_pw_conv_tma_set_enabled(int(bool(enabled)))


def set_bn_resident(enabled):
    """Channel-resident one-launch BatchNorm passes on small maps on (default) / off -- for A/B measurements and tests."""
    lib().rb_bn_set_resident(int(bool(enabled)))


def last_impl():
    return int(lib().rb_last_impl())
