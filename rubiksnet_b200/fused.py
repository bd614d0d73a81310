"""Fused building blocks of RubiksShiftBlock (rubiksnet/backbone.py:74-135) on CUDA:

  bn_act(x, bn, relu)            BatchNorm2d (+ReLU) in two streaming passes of librubiks_b200 (bn_relu.cu)
  conv1x1(x, weight, residual)   the block's 1x1 convolutions.  bf16 activations: librubiks_b200's tcgen05 GEMM
                                 kernels (pw_conv.cu) for forward, input gradient and weight gradient, the residual
                                 add `out += shortcut` in the epilogue.  fp32 activations: NCHW batched GEMMs on cuBLAS.
  rubiks_block(block, x)         a whole identity-shortcut RubiksShiftBlock (rubiks3d variant, bf16) as ONE autograd
                                 Function with a hand-scheduled kernel sequence:
                                   fwd  bn1 stats | conv2 with bn1+relu folded into its operand producer |
                                        bn2 stats+apply | 3D shift + conv3 + residual in one tensor-core launch
                                   bwd  conv3 dgrad | conv3 wgrad (shift recomputed in the producer) | shift backward |
                                        bn2 backward | conv2 dgrad | conv2 wgrad | bn1 backward (+ shortcut gradient)

All are autograd Functions over [N*T, C, H, W] activations; parameters stay fp32 (activations may be bf16).
"""
import threading

import torch
import torch.nn as nn

from . import _lib, ops
from .rubiksnet_cuda import _on_device

__all__ = ["bn_act", "conv1x1", "se_gate", "stem_conv", "rubiks_block", "rubiks_block_supported", "rubiks_down_block",
           "rubiks_down_block_supported", "rubiks_aq_block", "rubiks_aq_block_supported"]


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
            with _lib.timed("bn_forward<stats+apply>" if training else "bn_forward<eval>", _lib.nbytes(x) * (2 if training else 1) + _lib.nbytes(y)):
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
        dy = _aligned(dy)
        need_x = ctx.needs_input_grad[0]
        need_w = weight is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dx = torch.empty_like(x) if need_x else None
        dgamma = torch.empty(c, dtype=torch.float32, device=x.device) if need_w else None
        dbeta = torch.empty(c, dtype=torch.float32, device=x.device) if need_w else None
        with _on_device(x.device):
            L = _lib.lib()
            nbytes = L.rb_bn_workspace_bytes(ni, c)
            ws = _lib.workspace(nbytes, x.device)
            with _lib.timed("bn_backward", 2 * _lib.nbytes(x, dy) + _lib.nbytes(dx)):
                _lib.check(L.rb_bn_act_backward(
                    _lib.ptr(x), _lib.ptr(dy), None, _lib.ptr(weight), _lib.ptr(mean_invstd), _lib.ptr(scale_bias),
                    _lib.ptr(dx), _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.dtype_code(x), ni, c, hw, int(training),
                    int(relu), _lib.ptr(ws), nbytes, _lib.stream_handle(x.device)))
        return dx, dgamma, dbeta, None, None, None, None, None, None


def _f32(t):
    return t if t is None or t.dtype == torch.float32 else t.float()


def _bn_momentum(bn):
    """Exponential-average factor of nn.BatchNorm2d: `momentum`, or the cumulative average 1/num_batches_tracked
    (read AFTER the increment of this step) when momentum is None."""
    if bn.momentum is not None:
        return float(bn.momentum)
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        return 1.0 / max(float(bn.num_batches_tracked.item()), 1.0)
    return 0.0


class _RunningStats:
    """fp32 views of a BatchNorm's running statistics for the kernels (which read / update `float *`).  16-bit buffers
    (model.half() / model.bfloat16()) are updated through fp32 copies that are written back on exit."""

    def __init__(self, bn):
        track = bn.track_running_stats and bn.running_mean is not None
        self.rm = bn.running_mean if track else None
        self.rv = bn.running_var if track else None
        self.copy_back = track and bn.training and (self.rm.dtype != torch.float32 or self.rv.dtype != torch.float32)
        self.mean, self.var = _f32(self.rm), _f32(self.rv)

    def commit(self):
        if self.copy_back:
            self.rm.copy_(self.mean)
            self.rv.copy_(self.var)


def _aligned(t):
    """The BN passes use 128-bit accesses: a contiguous view with a storage offset that is not 16-byte aligned is copied."""
    t = t.contiguous()
    return t if t.data_ptr() % 16 == 0 else t.clone(memory_format=torch.contiguous_format)


class _StepCounters(threading.local):
    """nn.BatchNorm2d counts its training batches in `num_batches_tracked`; one `add_` per layer is 103 one-thread launches
    in a RubiksNet-Large step.  A network-level forward (RubiksNetBackbone) bumps every counter with ONE multi-tensor add
    and the per-layer code then skips the modules it covered."""
    done = None


_STEP_COUNTERS = _StepCounters()


def begin_step_counters(owner):
    bns = [m for m in owner.modules() if isinstance(m, nn.BatchNorm2d) and m.training and m.track_running_stats
           and m.num_batches_tracked is not None and m.num_batches_tracked.is_cuda]
    if bns:
        torch._foreach_add_([m.num_batches_tracked for m in bns], 1)
    _STEP_COUNTERS.done = {id(m) for m in bns}


def end_step_counters():
    _STEP_COUNTERS.done = None


def _count_batch(bn):
    if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        done = _STEP_COUNTERS.done
        if done is None or id(bn) not in done:
            bn.num_batches_tracked.add_(1)


def bn_act(x, bn, relu=True):
    """relu(bn(x)) with torch.nn.BatchNorm2d semantics (batch statistics + running-stat update in training
    mode, running statistics in eval mode) for a CUDA NCHW tensor."""
    training = bn.training or bn.running_mean is None
    _count_batch(bn)
    rs = _RunningStats(bn)
    out = _BNAct.apply(_aligned(x), _f32(bn.weight), _f32(bn.bias), rs.mean, rs.var, training, _bn_momentum(bn), bn.eps, relu)
    rs.commit()
    return out


# True: 1x1 convolutions whose geometry has an image path run on the second-generation TMA kernel (csrc/pw_conv2.cu);
# False: always the first-generation k_pw_conv.  Same arithmetic either way (tests/test_gpu_pwconv2.py).
USE_IMAGE_KERNEL = True


def _pack_weight(weight, ni, hw):
    """bf16 operands of a fp32 conv weight [N,K,1,1] for the forward and the input-gradient GEMM: packed images for the
    TMA kernel where that is the faster schedule, else the plain [N,K] / [K,N] copies.  Inside a forward pass of a module
    that called begin_step_pack() both kinds come out of ONE launch per step."""
    n, k = weight.shape[0], weight.shape[1]
    # the optional single-launch shift+conv3 / epilogue-statistics schedules exist on the first-generation kernel only
    image = (USE_IMAGE_KERNEL and not FUSE_SHIFT_CONV3 and not EPILOGUE_BN_STATS
             and ops.pw_image_supported(ni, k, n, hw, True, preferred=True)
             and ops.pw_image_supported(ni, n, k, hw, False, preferred=True))
    pre = _STEP_PACK.lookup(weight, image)
    if pre is not None:
        return pre
    _STEP_PACK.note(weight, image)
    return ops.pw_weight_images(weight) if image else ops.pw_weight_pack(weight)


class _StepPack:
    """All packed weight operands of a model in ONE launch per kind and step (rb_pw_weight_pack_multi /
    rb_pw_weight_image_pack_multi) instead of one or two small launches per conv and step (106 `pw_weight_pack` launches,
    2.4 ms of an eager RubiksNet-Large step in round 1).

    The first forward pass of an `owner` module packs every weight on its own and records which weights asked for which
    kind; from the second pass on `begin(owner)` fills all of them with one kernel per kind into a persistent buffer and
    `_pack_weight` hands out views of it.  The operands of step i are overwritten by `begin` of step i+1, i.e. after the
    backward pass of step i has been enqueued on the same stream."""

    def __init__(self):
        self.current = None      # {(id(weight), image?): (fwd operand, bwd operand)} of the running forward pass
        self.collecting = None   # [(weight, image?)] seen in a recording pass

    def lookup(self, weight, image):
        if self.current is None:
            return None
        return self.current.get((id(weight), image))

    def note(self, weight, image):
        if self.collecting is not None and all(not (w is weight and im == image) for w, im in self.collecting):
            self.collecting.append((weight, image))

    def begin(self, owner):
        plan = owner.__dict__.get("_rb_pack_plan")
        self.current, self.collecting = None, None
        if plan is not None and not plan.valid():
            plan = None
            owner.__dict__.pop("_rb_pack_plan", None)
        if plan is None:
            self.collecting = []
            return
        plan.launch()
        self.current = plan.operands

    def end(self, owner):
        if self.collecting:
            owner.__dict__["_rb_pack_plan"] = _PackPlan(self.collecting)
        self.current, self.collecting = None, None


class _PackPlan:
    def __init__(self, entries):
        import struct
        self.entries = list(entries)
        self.ptrs = [w.data_ptr() for w, _ in self.entries]
        dev = self.entries[0][0].device
        L = _lib.lib()
        sizes = []
        for w, image in self.entries:
            n, k = w.shape[0], w.shape[1]
            sizes.append((L.rb_pw_weight_image_bytes(n, k), L.rb_pw_weight_image_bytes(k, n)) if image else (2 * n * k, 2 * n * k))
        al = lambda v: (v + 255) // 256 * 256  # noqa: E731
        self.buffer = torch.empty(sum(al(a) + al(b) for a, b in sizes), dtype=torch.uint8, device=dev)
        self.operands = {}
        tables = {True: b"", False: b""}
        off = 0
        for (w, image), (a, b) in zip(self.entries, sizes):
            n, k = w.shape[0], w.shape[1]
            fwd = self.buffer[off:off + a]
            off += al(a)
            bwd = self.buffer[off:off + b]
            off += al(b)
            if image:
                self.operands[(id(w), True)] = (ops.WeightImage(fwd, n, k), ops.WeightImage(bwd, k, n))
            else:
                self.operands[(id(w), False)] = (fwd.view(torch.bfloat16).view(n, k), bwd.view(torch.bfloat16).view(k, n))
            tables[image] += struct.pack("<QQQii", w.data_ptr(), fwd.data_ptr(), bwd.data_ptr(), n, k)  # rb_pw_pack_item_t
        self.tables = {im: (torch.frombuffer(bytearray(t), dtype=torch.uint8).to(dev), len(t) // 32) for im, t in tables.items() if t}
        self.device = dev

    def valid(self):
        return all(w.data_ptr() == p and w.dtype == torch.float32 and w.is_contiguous() and w.device == self.device
                   for (w, _), p in zip(self.entries, self.ptrs))

    def launch(self):
        L = _lib.lib()
        with _on_device(self.device):
            with _lib.timed("pw_weight_pack", 0):
                for image, (table, count) in self.tables.items():
                    fn = L.rb_pw_weight_image_pack_multi if image else L.rb_pw_weight_pack_multi
                    _lib.check(fn(_lib.ptr(table), count, _lib.stream_handle(self.device)))


_STEP_PACK = _StepPack()


def begin_step_pack(owner):
    """Call at the start of a forward pass of `owner` (RubiksNetBackbone does): one launch packs every weight image."""
    _STEP_PACK.begin(owner)


def end_step_pack(owner):
    _STEP_PACK.end(owner)


def _wsave(w):
    """(tensor to save for backward, metadata) of a packed weight operand."""
    if isinstance(w, ops.WeightImage):
        return w.image, (w.rows, w.contraction)
    return w, None


def _wload(t, meta):
    return t if meta is None or t is None else ops.WeightImage(t, *meta)


class _Conv1x1TC(torch.autograd.Function):
    """bf16 activations: tcgen05 kernels of librubiks_b200 (fp32 master weight read directly by the kernels)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, residual):
        x = x.contiguous()
        if residual is not None:
            residual = residual.contiguous()
        ctx.has_res = residual is not None
        if weight.dtype == torch.float32 and weight.is_contiguous():
            w_nk, w_kn = _pack_weight(weight, x.shape[0], x.shape[2] * x.shape[3])
        else:
            w_nk, w_kn = weight, None
        w_kn_t, ctx.wmeta = _wsave(w_kn)
        ctx.save_for_backward(x, weight, w_kn_t)
        return ops.pw_conv(x, w_nk, residual=residual, resident=w_kn is not None)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        x, weight, w_kn = ctx.saved_tensors
        w_kn = _wload(w_kn, ctx.wmeta)
        g = g.contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = (ops.pw_conv(g, w_kn, name="pw_conv<dgrad>", resident=True) if w_kn is not None
                  else ops.pw_conv(g, weight, transposed=True, name="pw_conv<dgrad>"))
        gw = None
        if ctx.needs_input_grad[1]:
            gw = ops.pw_conv_wgrad(g, x).view(weight.shape).to(weight.dtype)
        gres = g if ctx.has_res and ctx.needs_input_grad[2] else None
        return gx, gw, gres


class _conv_tf32:
    """fp32 1x1 convolutions follow torch.backends.cudnn.allow_tf32 -- the switch that governs nn.Conv2d, which is what the
    reference runs here (default True: TF32 tensor cores) -- although they are expressed as batched GEMMs."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(torch.backends.cudnn.allow_tf32)

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


class _Conv1x1(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, residual):
        with _conv_tf32():
            return _Conv1x1._forward(ctx, x, weight, residual)

    @staticmethod
    def _forward(ctx, x, weight, residual):
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
        with _conv_tf32():
            return _Conv1x1._backward(ctx, g)

    @staticmethod
    def _backward(ctx, g):
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
    if x.dtype == torch.bfloat16 and weight.dtype in (torch.float32, torch.bfloat16):
        return _Conv1x1TC.apply(x, weight, residual)
    return _Conv1x1.apply(x, weight, residual)


# ------------------------------------------------------------------------------------- first layer (conv1)

class _StemConv(torch.autograd.Function):
    """conv1 (3x3, stride 2, padding 1, no bias; rubiksnet/backbone.py:148-149) on the library's own kernels: one im2col
    launch builds the bf16 patch matrix [NI, 32, Ho, Wo] (27 taps + 5 zero rows) straight from the fp32 clips, the
    convolution is the tcgen05 1x1 GEMM on it and the weight gradient the tcgen05 weight-gradient kernel on the saved
    patches -- instead of a cast, cuDNN's implicit GEMM and two NCHW<->NHWC transposes per direction (2.5 ms of a
    RubiksNet-Large step at 32 clips).  The input gets no gradient (clips are data)."""

    @staticmethod
    def forward(ctx, x, weight, stride):
        cout, cin = weight.shape[0], weight.shape[1]
        t = 9 * cin
        tpad = (t + 15) // 16 * 16
        cols = ops.im2col3x3(x.contiguous(), stride, tpad, torch.bfloat16)
        w = torch.nn.functional.pad(weight.detach().reshape(cout, t).float(), (0, tpad - t)).contiguous()
        ctx.save_for_backward(cols)
        ctx.wshape, ctx.wdtype, ctx.t = weight.shape, weight.dtype, t
        return ops.pw_conv(cols, w, name="pw_conv<conv1>")

    @staticmethod
    def backward(ctx, g):
        (cols,) = ctx.saved_tensors
        gw = None
        if ctx.needs_input_grad[1]:
            gw = ops.pw_conv_wgrad(g.contiguous(), cols, name="pw_conv_wgrad<conv1>")[:, :ctx.t].reshape(ctx.wshape).to(ctx.wdtype)
        return None, gw, None


def _stem_geometry_ok(conv, x):
    return (isinstance(conv, nn.Conv2d) and conv.kernel_size == (3, 3) and conv.padding == (1, 1) and conv.dilation == (1, 1)
            and conv.stride[0] == conv.stride[1] and conv.stride[0] in (1, 2) and conv.groups == 1 and conv.bias is None and conv.padding_mode == "zeros"
            and x.is_cuda and x.dim() == 4 and not x.requires_grad and ((x.shape[3] - 1) // conv.stride[0] + 1) % 8 == 0)


def stem_conv(conv, x):
    """conv(x) for the network's first layer.  bf16 autocast (training or inference): im2col + tcgen05 GEMM (+ tcgen05
    weight gradient); fp32 inference with TF32 convolutions allowed: im2col + the kind::tf32 GEMM; anything else (fp64,
    fp32 training, an input that needs a gradient, odd geometries): the nn.Conv2d itself."""
    if not _stem_geometry_ok(conv, x):
        return conv(x)
    if (torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and x.dtype in (torch.float32, torch.bfloat16) and conv.weight.dtype == torch.float32):
        with torch.autocast(device_type="cuda", enabled=False):
            return _StemConv.apply(x, conv.weight, conv.stride[0])
    if (x.dtype == torch.float32 and conv.weight.dtype == torch.float32 and not torch.is_grad_enabled()
            and not torch.is_autocast_enabled("cuda") and torch.backends.cudnn.allow_tf32):
        t = 9 * conv.weight.shape[1]
        cols = ops.im2col3x3(x.contiguous(), conv.stride[0], t, torch.float32)
        return ops.pw_conv_f32(cols, conv.weight.reshape(conv.weight.shape[0], t), resident=True, name="pw_conv_tf32<conv1>")
    return conv(x)


# ------------------------------------------------------------------------------------- squeeze-and-excitation

def _plane_reduce(a, b, scale):
    ni, c = a.shape[0], a.shape[1]
    hw = a.numel() // max(ni * c, 1)
    out = torch.empty(ni, c, dtype=torch.float32, device=a.device)
    with _on_device(a.device), _lib.timed("se_plane_reduce", _lib.nbytes(a, b)):
        _lib.check(_lib.lib().rb_plane_reduce(_lib.ptr(a), _lib.ptr(b), _lib.ptr(out), _lib.dtype_code(a), ni * c, hw, float(scale),
                                              _lib.stream_handle(a.device)))
    return out


def _plane_scale(a, s, t):
    assert s.dtype == torch.float32 and s.is_contiguous() and (t is None or (t.dtype == torch.float32 and t.is_contiguous()))
    ni, c = a.shape[0], a.shape[1]
    hw = a.numel() // max(ni * c, 1)
    out = torch.empty_like(a)
    with _on_device(a.device), _lib.timed("se_plane_scale", _lib.nbytes(a, out)):
        _lib.check(_lib.lib().rb_plane_scale(_lib.ptr(a), _lib.ptr(s), _lib.ptr(t), _lib.ptr(out), _lib.dtype_code(a), ni * c, hw,
                                             _lib.stream_handle(a.device)))
    return out


class _SEGate(torch.autograd.Function):
    """y = x * sigmoid(W2 relu(W1 avgpool(x)))  (SELayer, rubiksnet/backbone.py:56-71).  The two passes over x (pool,
    rescale) and the two passes of the backward (sum of g*x per plane, dx = g*gate + dpool/HW) are librubiks_b200 kernels;
    the gate MLP and its gradients act on [NI, C] / [NI, C/r] matrices."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, w1, w2):
        x = x.contiguous()
        hw = x.shape[2] * x.shape[3]
        with torch.autocast(device_type="cuda", enabled=False):  # the gate MLP stays in fp32 (the kernels read float gates)
            pooled = _plane_reduce(x, None, 1.0 / hw)                      # [NI, C] fp32
            w1f, w2f = w1.float(), w2.float()
            hidden = torch.relu(pooled @ w1f.t())                          # [NI, C/r]
            gate = torch.sigmoid(hidden @ w2f.t()).contiguous()            # [NI, C]
        ctx.save_for_backward(x, pooled, hidden, gate, w1f, w2f)
        ctx.wdtypes = (w1.dtype, w2.dtype)
        return _plane_scale(x, gate, None)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        x, pooled, hidden, gate, w1f, w2f = ctx.saved_tensors
        g = g.contiguous()
        hw = x.shape[2] * x.shape[3]
        with torch.autocast(device_type="cuda", enabled=False):
            dgate = _plane_reduce(g, x, 1.0)                               # sum_p g * x
            dz2 = dgate * gate * (1.0 - gate)
            dw2 = dz2.t() @ hidden
            dz1 = (dz2 @ w2f) * (hidden > 0).to(dz2.dtype)
            dw1 = dz1.t() @ pooled
            dpool = ((dz1 @ w1f) * (1.0 / hw)).contiguous()
        dx = _plane_scale(g, gate, dpool) if ctx.needs_input_grad[0] else None
        return dx, dw1.to(ctx.wdtypes[0]), dw2.to(ctx.wdtypes[1])


def se_gate(x, se):
    """SELayer forward on a CUDA NCHW tensor in fp32 / fp16 / bf16 (se.fc = Linear, ReLU, Linear, Sigmoid; no biases)."""
    return _SEGate.apply(x, se.fc[0].weight, se.fc[2].weight)


# ------------------------------------------------------------------------------------- eval-mode block (fp32)

def _eval_bn_ok(bn):
    return (not bn.training and bn.running_mean is not None and bn.running_mean.dtype == torch.float32
            and bn.running_var.dtype == torch.float32 and bn.weight is not None and bn.weight.dtype == torch.float32)


def eval_block_supported(block, x):
    """fp32 inference (no autograd, BatchNorms in eval mode) with TF32 convolutions allowed -- the switch nn.Conv2d follows
    in the reference, default True: the block runs on the tcgen05 kind::tf32 kernel with both BatchNorm+ReLU pairs folded
    into the two GEMMs.  torch.backends.cudnn.allow_tf32 = False keeps full-fp32 GEMMs (cuBLAS)."""
    if not (x.is_cuda and x.dtype == torch.float32 and not torch.is_grad_enabled() and torch.backends.cudnn.allow_tf32):
        return False
    conv2 = block.conv2[1] if isinstance(block.conv2, nn.Sequential) else block.conv2
    convs = [conv2, block.conv3] + ([] if isinstance(block.shortcut, nn.Identity) else [block.shortcut])
    return (_eval_bn_ok(block.bn1) and _eval_bn_ok(block.bn2)
            and all(c.weight.dtype == torch.float32 and c.weight.is_contiguous() for c in convs))


def _eval_coeffs(bn):
    """(scale, bias) [C, 2] of an eval-mode BatchNorm (rb_bn_act_forward, training = 0, no apply pass: one tiny launch).
    Cached on the module and keyed by the version counters of its parameters / buffers, so steady-state inference launches
    nothing; while a CUDA graph is being captured the kernel is always recorded, so that replays follow the buffers."""
    tensors = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t.data_ptr(), t._version) for t in tensors) + (bn.eps,)
    capturing = torch.cuda.is_current_stream_capturing()
    hit = bn.__dict__.get("_rb_eval_sb")
    if hit is not None and hit[0] == key and not capturing:
        return hit[1]
    like = torch.empty(1, bn.num_features, 1, 4, dtype=torch.float32, device=bn.weight.device)  # shape carrier, never read
    _, _, sb = ops.bn_forward(like, bn.weight, bn.bias, bn.running_mean, bn.running_var, False, 0.0, bn.eps, relu=True, apply=False)
    if not capturing:
        bn.__dict__["_rb_eval_sb"] = (key, sb)
    return sb


def eval_block(block, x):
    """RubiksShiftBlock.forward (rubiksnet/backbone.py:109-135) in eval mode as three launches for an identity-shortcut
    block: conv2 [bn1+relu in the operand producer, bn2+relu in the epilogue] -> shift -> conv3 [+ shortcut in the epilogue]."""
    x = _aligned(x)
    sb1, sb2 = _eval_coeffs(block.bn1), _eval_coeffs(block.bn2)
    aq = isinstance(block.conv2, nn.Sequential)
    identity = isinstance(block.shortcut, nn.Identity)
    w2 = (block.conv2[1] if aq else block.conv2).weight
    if identity and not aq:
        shortcut = x
        a2 = ops.pw_conv_f32(x, w2, in_scale_bias=sb1, out_scale_bias=sb2, relu=True, resident=True)
    else:
        o = ops.bn_apply(x, sb1, relu=True)
        if identity:
            shortcut = x
        else:
            st = block.shortcut.stride[0]
            shortcut = ops.pw_conv_f32(o if st == 1 else o[:, :, ::st, ::st].contiguous(), block.shortcut.weight, resident=True)
        a2 = ops.pw_conv_f32(block.conv2[0](o).contiguous() if aq else o, w2, out_scale_bias=sb2, relu=True, resident=True)
    s3 = block.as3(a2)
    if block.se is not None:
        s3 = se_gate(s3, block.se)
    return ops.pw_conv_f32(s3.contiguous(), block.conv3.weight, residual=shortcut, resident=True)


# ------------------------------------------------------------------------------------- whole block

def _bn_cfg(bn):
    """(training, momentum, eps, running_mean, running_var) for the whole-block Functions; they are only used when the
    buffers are fp32 (rubiks_block_supported checks), so the kernels may update them in place."""
    training = bn.training or bn.running_mean is None
    track = bn.track_running_stats and bn.running_mean is not None
    return training, _bn_momentum(bn), bn.eps, (bn.running_mean if track else None), (bn.running_var if track else None)


# True: as3 -> conv3 -> += shortcut run as ONE launch (the shift is the operand producer of the tensor-core GEMM and is
# recomputed by the conv3 weight-gradient kernel; the shifted tensor never exists in HBM).  False: the stand-alone strip
# shift kernel writes the shifted tensor, which conv3 and its weight gradient then read.  Measured on B200 at 32 clips
# (tools/bench_pw.py, profiles/): the per-element gather inside the GEMM producer costs more issue slots than the strip
# kernel's shared-memory staging saves in HBM traffic, so the two-launch schedule is the faster one on every
# RubiksNet-Large geometry today and is the default.
FUSE_SHIFT_CONV3 = False

# True: the per-channel sums BatchNorm needs are reduced in the epilogue of the GEMM that produces the tensor (conv2 ->
# bn2 inside a block, conv3 + shortcut -> bn1 of the next block) instead of by a separate pass over it.  Parity-tested
# (tests/test_gpu_pwconv.py::test_pw_conv_epilogue_bn_statistics) but OFF by default: measured on B200 at 32 clips
# (gpurun_out/s27, round 1) it removes 2.6 ms of statistics passes per step and adds 3.0 ms to the two GEMMs, whose
# epilogue warps -- not HBM -- are what bounds them (731 vs 756 clips/s).
EPILOGUE_BN_STATS = False


def _shift3d_forward(x, shift, frames, stride=1):
    from .shiftlib.rubiks3d.primitive import rubiks_shift_3d_forward
    nt, c, h, w = x.shape
    out = rubiks_shift_3d_forward(x.view(nt // frames, frames, c, h, w), shift, (1, stride, stride), 0)
    return out.view(nt, c, out.shape[-2], out.shape[-1])


class _RubiksBlockFn(torch.autograd.Function):
    """Identity-shortcut RubiksShiftBlock with a 3D shift (backbone.py:109-135 + models.py:128-145), bf16."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, g1, b1, w2, g2, b2, shift, w3, bn1, bn2, frames, normalize_grad, normalize_t_factor, x_stats):
        x = x.contiguous()
        tr1, mom1, eps1, rm1, rv1 = _bn_cfg(bn1)
        tr2, mom2, eps2, rm2, rv2 = _bn_cfg(bn2)
        count = x.shape[0] * x.shape[2] * x.shape[3]
        # conv weights: rounded to bf16 once per step, in both orientations (forward / input gradient)
        hw = x.shape[2] * x.shape[3]
        w2_nk, w2_kn = _pack_weight(w2, x.shape[0], hw)
        w3_nk, w3_kn = _pack_weight(w3, x.shape[0], hw)
        # BatchNorm statistics come out of the epilogue of the GEMM that produced the tensor (x: the previous block's
        # conv3, handed over as x_stats; y2: conv2 below) -- no separate reduction pass in training mode
        if x_stats is not None and tr1:
            mi1, sb1 = ops.bn_finalize(x_stats, count, g1, b1, rm1, rv1, mom1, eps1)
        else:
            _, mi1, sb1 = ops.bn_forward(x, g1, b1, rm1, rv1, tr1, mom1, eps1, relu=True, apply=False)
        # ... always where the epilogue statistics are free (tensor-map schedule on 16-byte-pitch maps: an epilogue thread owns
        # a channel row), everywhere with EPILOGUE_BN_STATS (slower than the separate pass on the other kernels)
        mid, cout = w2.shape[0], w3.shape[0]
        if tr2 and not isinstance(w2_nk, ops.WeightImage) and (EPILOGUE_BN_STATS or ops.pw_stats_preferred(x.shape[0], x.shape[1], mid, hw)):
            y2, st2 = ops.pw_conv(x, w2_nk, in_scale_bias=sb1, name="pw_conv<bn+relu>", stats=True, resident=True)
            mi2, sb2 = ops.bn_finalize(st2, count, g2, b2, rm2, rv2, mom2, eps2)
            a2 = ops.bn_apply(y2, sb2, relu=True)
        else:
            y2 = ops.pw_conv(x, w2_nk, in_scale_bias=sb1, name="pw_conv<bn+relu>", resident=True)
            a2, mi2, sb2 = ops.bn_forward(y2, g2, b2, rm2, rv2, tr2, mom2, eps2, relu=True, apply=True)
        out_stats = None
        if FUSE_SHIFT_CONV3 and not isinstance(w3_nk, ops.WeightImage):
            s3 = None
            out = ops.shift3d_pw_conv(a2, shift, w3_nk, x, frames)
        else:
            s3 = _shift3d_forward(a2, shift, frames)
            if not isinstance(w3_nk, ops.WeightImage) and (EPILOGUE_BN_STATS or ops.pw_stats_preferred(x.shape[0], mid, cout, hw)):
                out, out_stats = ops.pw_conv(s3, w3_nk, residual=x, name="pw_conv<+residual>", stats=True, resident=True)
            else:
                out = ops.pw_conv(s3, w3_nk, residual=x, name="pw_conv<+residual>", resident=True)
        w2_kn_t, m2 = _wsave(w2_kn)
        w3_kn_t, m3 = _wsave(w3_kn)
        ctx.save_for_backward(x, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, shift, w2_kn_t, w3_kn_t)
        ctx.cfg = (tr1, tr2, frames, normalize_grad, normalize_t_factor)
        ctx.wmeta = (m2, m3)
        if out_stats is None:
            return out, None, 0
        ctx.mark_non_differentiable(out_stats[0])
        return out, out_stats[0], out_stats[1]

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g, _g_stats=None, _g_splits=None):
        x, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, shift, w2_kn, w3_kn = ctx.saved_tensors
        tr1, tr2, frames, normalize_grad, normalize_t_factor = ctx.cfg
        w2_kn, w3_kn = _wload(w2_kn, ctx.wmeta[0]), _wload(w3_kn, ctx.wmeta[1])
        g = g.contiguous()
        need = ctx.needs_input_grad
        gs = ops.pw_conv(g, w3_kn, name="pw_conv<dgrad>", resident=True)
        gw3 = None
        if need[7]:
            gw3 = (ops.shift3d_pw_conv_wgrad(g, a2, shift, frames) if s3 is None else ops.pw_conv_wgrad(g, s3)).view(w3.shape)
        del s3
        ga2, gshift = ops.shift3d_backward(a2, shift, gs, frames, normalize_grad, normalize_t_factor, need_shift=need[6])
        del gs
        gy2, dg2, db2 = ops.bn_backward(y2, ga2, None, g2, mi2, sb2, tr2, relu=True)
        del ga2
        go = ops.pw_conv(gy2, w2_kn, name="pw_conv<dgrad>", resident=True)
        gw2 = ops.pw_conv_wgrad(gy2, x, in_scale_bias=sb1, name="pw_conv_wgrad<bn+relu>").view(w2.shape) if need[3] else None
        del gy2
        gx, dg1, db1 = ops.bn_backward(x, go, g, g1, mi1, sb1, tr1, relu=True, need_dx=need[0])
        return gx, dg1, db1, gw2, dg2, db2, gshift, gw3, None, None, None, None, None, None


class _RubiksDownBlockFn(torch.autograd.Function):
    """Down-sampling RubiksShiftBlock (first block of layer1..4: 3D shift with stride (1,2,2), 1x1 shortcut conv with stride 2,
    rubiksnet/backbone.py:104-105,109-135) as one Function, bf16.  relu(bn1(x)) is never materialised: conv2 applies it in
    its operand producer, and the shortcut conv -- which only sees every other row / column -- gets it from a BN-apply pass
    over the sub-sampled quarter of x.  In the backward pass the shortcut's input gradient is added into the sub-sampled
    positions of conv2's input gradient in place (one strided pass over a quarter of the elements) instead of autograd's
    zero-fill + scatter + full-size add."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, g1, b1, w2, g2, b2, shift, w3, wsc, bn1, bn2, frames, normalize_grad, normalize_t_factor):
        x = x.contiguous()
        tr1, mom1, eps1, rm1, rv1 = _bn_cfg(bn1)
        tr2, mom2, eps2, rm2, rv2 = _bn_cfg(bn2)
        ni, hw = x.shape[0], x.shape[2] * x.shape[3]
        w2_nk, w2_kn = _pack_weight(w2, ni, hw)
        w3_nk, w3_kn = _pack_weight(w3, ni, hw // 4)
        wsc_nk, wsc_kn = _pack_weight(wsc, ni, hw // 4)
        _, mi1, sb1 = ops.bn_forward(x, g1, b1, rm1, rv1, tr1, mom1, eps1, relu=True, apply=False)
        o_sub = ops.bn_apply(x[:, :, ::2, ::2].contiguous(), sb1, relu=True)  # relu(bn1(x)) where the shortcut looks
        sc = ops.pw_conv(o_sub, wsc_nk, name="pw_conv", resident=True)
        y2 = ops.pw_conv(x, w2_nk, in_scale_bias=sb1, name="pw_conv<bn+relu>", resident=True)
        a2, mi2, sb2 = ops.bn_forward(y2, g2, b2, rm2, rv2, tr2, mom2, eps2, relu=True, apply=True)
        s3 = _shift3d_forward(a2, shift, frames, stride=2)
        out = ops.pw_conv(s3, w3_nk, residual=sc, name="pw_conv<+residual>", resident=True)
        del sc
        saved_w = [t for w in (w2_kn, w3_kn, wsc_kn) for t in (_wsave(w)[0],)]
        ctx.wmeta = tuple(_wsave(w)[1] for w in (w2_kn, w3_kn, wsc_kn))
        ctx.save_for_backward(x, o_sub, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, wsc, shift, *saved_w)
        ctx.cfg = (tr1, tr2, frames, normalize_grad, normalize_t_factor)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        x, o_sub, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, wsc, shift, w2_kn, w3_kn, wsc_kn = ctx.saved_tensors
        tr1, tr2, frames, normalize_grad, normalize_t_factor = ctx.cfg
        w2_kn, w3_kn, wsc_kn = (_wload(t, m) for t, m in zip((w2_kn, w3_kn, wsc_kn), ctx.wmeta))
        g = g.contiguous()
        need = ctx.needs_input_grad
        gs = ops.pw_conv(g, w3_kn, name="pw_conv<dgrad>", resident=True)
        gw3 = ops.pw_conv_wgrad(g, s3).view(w3.shape) if need[7] else None
        del s3
        ga2, gshift = ops.shift3d_backward(a2, shift, gs, frames, normalize_grad, normalize_t_factor, need_shift=need[6], stride=2)
        del gs
        gy2, dg2, db2 = ops.bn_backward(y2, ga2, None, g2, mi2, sb2, tr2, relu=True)
        del ga2
        go = ops.pw_conv(gy2, w2_kn, name="pw_conv<dgrad>", resident=True)
        gw2 = ops.pw_conv_wgrad(gy2, x, in_scale_bias=sb1, name="pw_conv_wgrad<bn+relu>").view(w2.shape) if need[3] else None
        del gy2
        # shortcut path: `out = conv3(...) + sc` hands g to the shortcut conv unchanged
        gwsc = ops.pw_conv_wgrad(g, o_sub).view(wsc.shape) if need[8] else None
        go[:, :, ::2, ::2] += ops.pw_conv(g, wsc_kn, name="pw_conv<dgrad>", resident=True)
        gx, dg1, db1 = ops.bn_backward(x, go, None, g1, mi1, sb1, tr1, relu=True, need_dx=need[0])
        return gx, dg1, db1, gw2, dg2, db2, gshift, gw3, gwsc, None, None, None, None, None


class _RubiksAQBlockFn(torch.autograd.Function):
    """Identity-shortcut block of the attention-quantized variant (rubiksnet/models.py:100-104: AttentionShift in front of
    conv2, the block keeps its 2D spatial shift) as one Function, bf16:
        fwd  bn1 stats | temporal mix of relu(bn1(x)) | conv2 | bn2 stats+apply | 2D shift | conv3 + residual
        bwd  conv3 dgrad / wgrad | 2D shift bwd | bn2 bwd | conv2 dgrad / wgrad | temporal mix bwd | bn1 bwd + shortcut gradient
    relu(bn1(x)) is never written: both temporal-mix kernels apply it to x as they load (rb_bn_attention_shift_*).
    `taps` is the [C, 3] softmax of the attention weights, computed by autograd outside (its gradient is returned)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, g1, b1, taps, w2, g2, b2, shift, w3, bn1, bn2, frames, normalize_grad):
        from .attention_shift import attention_mix_forward
        from .shiftlib.rubiks2d.primitive import rubiks2d_forward
        x = x.contiguous()
        taps = taps.contiguous().float()
        tr1, mom1, eps1, rm1, rv1 = _bn_cfg(bn1)
        tr2, mom2, eps2, rm2, rv2 = _bn_cfg(bn2)
        hw = x.shape[2] * x.shape[3]
        w2_nk, w2_kn = _pack_weight(w2, x.shape[0], hw)
        w3_nk, w3_kn = _pack_weight(w3, x.shape[0], hw)
        _, mi1, sb1 = ops.bn_forward(x, g1, b1, rm1, rv1, tr1, mom1, eps1, relu=True, apply=False)
        att = attention_mix_forward(x, taps, frames, in_scale_bias=sb1)
        y2 = ops.pw_conv(att, w2_nk, name="pw_conv", resident=True)
        a2, mi2, sb2 = ops.bn_forward(y2, g2, b2, rm2, rv2, tr2, mom2, eps2, relu=True, apply=True)
        s3 = rubiks2d_forward(a2, shift, 1, 0)
        out = ops.pw_conv(s3, w3_nk, residual=x, name="pw_conv<+residual>", resident=True)
        w2_kn_t, m2 = _wsave(w2_kn)
        w3_kn_t, m3 = _wsave(w3_kn)
        ctx.save_for_backward(x, att, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, shift, taps, w2_kn_t, w3_kn_t)
        ctx.cfg = (tr1, tr2, frames, normalize_grad)
        ctx.wmeta = (m2, m3)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        from .attention_shift import attention_mix_backward
        from .shiftlib.rubiks2d.primitive import rubiks2d_backward
        x, att, y2, a2, s3, mi1, sb1, mi2, sb2, g1, g2, w2, w3, shift, taps, w2_kn, w3_kn = ctx.saved_tensors
        tr1, tr2, frames, normalize_grad = ctx.cfg
        w2_kn, w3_kn = _wload(w2_kn, ctx.wmeta[0]), _wload(w3_kn, ctx.wmeta[1])
        g = g.contiguous()
        need = ctx.needs_input_grad
        gs = ops.pw_conv(g, w3_kn, name="pw_conv<dgrad>", resident=True)
        gw3 = ops.pw_conv_wgrad(g, s3).view(w3.shape) if need[8] else None
        del s3
        ga2, gshift = rubiks2d_backward(gs, a2, shift, 1, 0, normalize_grad=normalize_grad, enable_shift_grad=bool(need[7]))
        del gs
        gy2, dg2, db2 = ops.bn_backward(y2, ga2, None, g2, mi2, sb2, tr2, relu=True)
        del ga2
        gatt = ops.pw_conv(gy2, w2_kn, name="pw_conv<dgrad>", resident=True)
        gw2 = ops.pw_conv_wgrad(gy2, att).view(w2.shape) if need[4] else None
        del gy2
        go, gtaps = attention_mix_backward(x, taps, gatt, frames, need_x=True, need_t=bool(need[3]), in_scale_bias=sb1)
        del gatt
        # the shortcut gradient g is added inside the bn1 backward pass (no separate add over the tensor)
        gx, dg1, db1 = ops.bn_backward(x, go, g, g1, mi1, sb1, tr1, relu=True, need_dx=need[0])
        return gx, dg1, db1, gtaps, gw2, dg2, db2, (gshift if need[7] else None), gw3, None, None, None, None


def rubiks_aq_block_supported(block, x):
    """True when `block` is an identity-shortcut block of the attention-quantized variant that can run as _RubiksAQBlockFn."""
    from .attention_shift import AttentionShift
    from .shiftlib import RubiksShift2D
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4 or x.data_ptr() % 16 != 0:
        return False
    if not isinstance(block.shortcut, nn.Identity) or block.se is not None or not isinstance(block.conv2, nn.Sequential):
        return False
    if len(block.conv2) != 2 or type(block.conv2[0]) is not AttentionShift or not isinstance(block.conv2[1], nn.Conv2d):
        return False
    att, as3 = block.conv2[0], block.as3
    if type(as3) is not RubiksShift2D or as3.stride != 1 or as3.padding != 0 or as3.quantize:
        return False
    if att.weight is None or x.shape[0] % att.n_segment != 0:
        return False
    params = (block.conv2[1].weight, block.conv3.weight, block.bn1.weight, block.bn2.weight, as3.shift, block.bn1.bias, block.bn2.bias)
    if any(p is None or p.dtype != torch.float32 for p in params):
        return False
    for bn in (block.bn1, block.bn2):
        if bn.track_running_stats and bn.running_mean is not None and (
                bn.running_mean.dtype != torch.float32 or bn.running_var.dtype != torch.float32):
            return False
    return True


def rubiks_aq_block(block, x):
    for bn in (block.bn1, block.bn2):
        _count_batch(bn)
    att = block.conv2[0]
    return _RubiksAQBlockFn.apply(
        x, block.bn1.weight, block.bn1.bias, att.taps().float(), block.conv2[1].weight, block.bn2.weight, block.bn2.bias,
        block.as3.shift, block.conv3.weight, block.bn1, block.bn2, att.n_segment, bool(block.as3.normalize_grad))


def _block_common_ok(block, x):
    """Conditions shared by the whole-block Functions: bf16 CUDA activations, no SE, plain conv2, a non-quantized 3D shift
    without padding behind _Rubiks3DWrap, fp32 parameters and BatchNorm buffers."""
    as3 = block.as3
    r3 = getattr(as3, "rubiks3d", None)
    if r3 is None or not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4:
        return False
    if block.se is not None or not isinstance(block.conv2, nn.Conv2d):
        return False
    if tuple(r3.padding) != (0, 0, 0) or r3.quantize:
        return False
    if not isinstance(r3.normalize_t_factor, (int, float)) or x.shape[0] % as3.n_segment != 0:
        return False
    params = [block.conv2.weight, block.conv3.weight, block.bn1.weight, block.bn2.weight, r3.shift, block.bn1.bias, block.bn2.bias]
    if not isinstance(block.shortcut, nn.Identity):
        params.append(block.shortcut.weight)
    if any(p is None or p.dtype != torch.float32 for p in params):
        return False
    for bn in (block.bn1, block.bn2):  # the kernels update the running statistics in place as float *
        if bn.track_running_stats and bn.running_mean is not None and (
                bn.running_mean.dtype != torch.float32 or bn.running_var.dtype != torch.float32):
            return False
    if x.data_ptr() % 16 != 0:
        return False
    return getattr(r3, "shift_function", None) is _default_shift_function()


def rubiks_down_block_supported(block, x):
    """True when `block` is a down-sampling block that can run as _RubiksDownBlockFn: stride-2 1x1 shortcut conv, 3D shift
    with stride (1,2,2), even map size."""
    if not _block_common_ok(block, x) or not isinstance(block.shortcut, nn.Conv2d):
        return False
    sc = block.shortcut
    if sc.kernel_size != (1, 1) or sc.stride != (2, 2) or sc.padding != (0, 0) or sc.bias is not None or sc.groups != 1:
        return False
    if tuple(block.as3.rubiks3d.stride) != (1, 2, 2) or x.shape[2] % 2 or x.shape[3] % 2:
        return False
    return True


def rubiks_down_block(block, x):
    for bn in (block.bn1, block.bn2):
        _count_batch(bn)
    r3 = block.as3.rubiks3d
    return _RubiksDownBlockFn.apply(
        x, block.bn1.weight, block.bn1.bias, block.conv2.weight, block.bn2.weight, block.bn2.bias, r3.shift, block.conv3.weight,
        block.shortcut.weight, block.bn1, block.bn2, block.as3.n_segment, bool(r3.normalize_grad), float(r3.normalize_t_factor))


def rubiks_block_supported(block, x):
    """True when `block` can run as the single fused Function: bf16 CUDA activations, identity shortcut, no SE,
    plain conv2, as3 = _Rubiks3DWrap around a stride-1 / pad-0 / non-quantized RubiksShift3D with fp32 parameters."""
    import torch.nn as nn
    as3 = block.as3
    r3 = getattr(as3, "rubiks3d", None)
    if r3 is None or not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4:
        return False
    if not isinstance(block.shortcut, nn.Identity) or block.se is not None or not isinstance(block.conv2, nn.Conv2d):
        return False
    if tuple(r3.stride) != (1, 1, 1) or tuple(r3.padding) != (0, 0, 0) or r3.quantize:
        return False
    if not isinstance(r3.normalize_t_factor, (int, float)) or x.shape[0] % as3.n_segment != 0:
        return False
    params = (block.conv2.weight, block.conv3.weight, block.bn1.weight, block.bn2.weight, r3.shift,
              block.bn1.bias, block.bn2.bias)
    if any(p is None or p.dtype != torch.float32 for p in params):
        return False
    for bn in (block.bn1, block.bn2):  # the kernels update the running statistics in place as float *
        if bn.track_running_stats and bn.running_mean is not None and (
                bn.running_mean.dtype != torch.float32 or bn.running_var.dtype != torch.float32):
            return False
    if x.data_ptr() % 16 != 0:
        return False
    return getattr(r3, "shift_function", None) is _default_shift_function()


def _default_shift_function():
    from .shiftlib.rubiks3d.primitive import rubiks_shift_3d
    return rubiks_shift_3d


def rubiks_block(block, x):
    for bn in (block.bn1, block.bn2):
        _count_batch(bn)
    r3 = block.as3.rubiks3d
    # statistics of x reduced by the conv3 epilogue of the block that produced it (attached to the tensor below)
    x_stats = getattr(x, "_rb_bn_stats", None)
    out, partial, splits = _RubiksBlockFn.apply(
        x, block.bn1.weight, block.bn1.bias, block.conv2.weight, block.bn2.weight, block.bn2.bias, r3.shift,
        block.conv3.weight, block.bn1, block.bn2, block.as3.n_segment, bool(r3.normalize_grad),
        float(r3.normalize_t_factor), x_stats)
    if partial is not None:
        out._rb_bn_stats = (partial, splits)
    return out
