"""GPU parity tests of the tcgen05 pointwise-conv kernels (pw_conv.cu) through the C ABI:
forward / input gradient / weight gradient against fp32 torch matmuls on the same bf16 inputs, the fused
3D-shift + conv3 launch against rb_shift3d_forward followed by the GEMM, and
the single-Function RubiksShiftBlock against the module graph.  Tolerance: 1e-2 relative (bf16, north_star)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200 import _lib, backbone, ops  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_forward  # noqa: E402

pytestmark = pytest.mark.gpu
BF = torch.bfloat16

# (NI, K, N, HW): every RubiksNet-Large / -Tiny block geometry class + ragged cases (HW not a multiple of 2/4/8,
# channels not a multiple of 8/16, N > 256 (two MMA column blocks), N split over CTAs (576), single pixel)
SHAPES = [(8, 72, 72, 12544), (8, 72, 72, 3136), (8, 144, 144, 784), (16, 288, 288, 196), (16, 576, 576, 49),
          (8, 72, 144, 784), (8, 288, 576, 49), (4, 54, 54, 3136), (4, 108, 216, 196), (4, 432, 432, 49),
          (3, 17, 23, 75), (2, 8, 300, 130), (1, 16, 16, 1), (5, 40, 24, 127)]


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1.0, b.double().abs().max().item())


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("wdtype", [torch.float32, BF])
def test_pw_conv_forward(shape, wdtype):
    ni, k, n, hw = shape
    torch.manual_seed(0)
    x = torch.randn(ni, k, hw, device="cuda").to(BF)
    w = (torch.randn(n, k, device="cuda") / k ** 0.5).to(wdtype)
    res = torch.randn(ni, n, hw, device="cuda").to(BF)
    ref = torch.matmul(w.to(BF).float(), x.float())
    out = ops.pw_conv(x, w)
    assert out.shape == (ni, n, hw) and out.dtype == BF
    assert _rel(out, ref) <= 1e-2
    out = ops.pw_conv(x, w, residual=res)
    assert _rel(out, ref + res.float()) <= 1e-2
    # input gradient = same kernel on the transposed weight buffer
    g = torch.randn(ni, n, hw, device="cuda").to(BF)
    gx = ops.pw_conv(g, w, transposed=True)
    assert gx.shape == x.shape
    assert _rel(gx, torch.matmul(w.to(BF).float().t(), g.float())) <= 1e-2


@pytest.mark.parametrize("geom", [(72, 112), (72, 56), (144, 28), (288, 14), (576, 7)])
def test_pw_conv_full_size_block_geometries(geom):
    """BASELINE C3 sizes (32 clips x 8 frames = 256 images, every RubiksNet-Large block geometry): forward with residual,
    BN+ReLU producer, input gradient on the packed transposed weight and weight gradient, checked on a sample of images
    against fp32 matmuls (weight gradient: all images), plus linearity of the forward in x (exact in the residual)."""
    c, h = geom
    ni, hw = 256, h * h
    torch.manual_seed(3)
    x = torch.randn(ni, c, hw, device="cuda").to(BF)
    res = torch.randn(ni, c, hw, device="cuda").to(BF)
    w = (torch.randn(c, c, device="cuda") / c ** 0.5).contiguous()
    w_nk, w_kn = ops.pw_weight_pack(w)
    sb = torch.stack([torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")], dim=1).contiguous()
    sample = [0, 1, 97, 254, 255]
    wf = w_nk.float()
    out = ops.pw_conv(x, w_nk, residual=res)
    assert _rel(out[sample], torch.matmul(wf, x[sample].float()) + res[sample].float()) <= 1e-2
    act = torch.relu(x.float() * sb[:, 0].view(1, c, 1) + sb[:, 1].view(1, c, 1)).to(BF)
    assert _rel(ops.pw_conv(x, w_nk, in_scale_bias=sb)[sample], torch.matmul(wf, act[sample].float())) <= 1e-2
    assert _rel(ops.pw_conv(x, w_kn)[sample], torch.matmul(wf.t(), x[sample].float())) <= 1e-2
    assert torch.equal(ops.pw_conv(x, w_nk), ops.pw_conv(x, w))  # packed weights = what the kernel rounds itself
    dw = ops.pw_conv_wgrad(res, x, in_scale_bias=sb)
    ref = torch.zeros(c, c, device="cuda", dtype=torch.float64)
    for i0 in range(0, ni, 32):
        ref += torch.einsum("inp,ikp->nk", res[i0:i0 + 32].double(), act[i0:i0 + 32].double())
    assert _rel(dw, ref) <= 1e-3
    zero = torch.zeros_like(x)
    assert torch.equal(ops.pw_conv(zero, w_nk, residual=res), res)  # W.0 + r == r exactly
    assert torch.count_nonzero(ops.pw_conv(zero, w_nk)).item() == 0


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("with_res", [False, True])
def test_pw_conv_epilogue_bn_statistics(shape, with_res):
    """rb_pw_conv_forward_stats: same output as rb_pw_conv_forward, and the per-channel sums reduced in the epilogue give
    -- through rb_bn_stats_finalize / rb_bn_apply_forward -- what torch's training-mode BatchNorm2d computes on that
    output (batch mean / biased variance, running statistics with the unbiased variance)."""
    ni, k, n, hw = shape
    torch.manual_seed(7)
    x = torch.randn(ni, k, hw, device="cuda").to(BF)
    w_nk, w_kn = ops.pw_weight_pack((torch.randn(n, k, device="cuda") / k ** 0.5).contiguous())
    res = torch.randn(ni, n, hw, device="cuda").to(BF) if with_res else None
    ref_out = ops.pw_conv(x, w_nk, residual=res)
    out, stats = ops.pw_conv(x, w_nk, residual=res, stats=True)
    assert torch.equal(out, ref_out)
    partial, splits = stats
    assert 1 <= splits <= 2 * 148
    sums = partial[:n * splits * 2].view(n, splits, 2).sum(1)
    o = out.double()
    assert _rel(sums[:, 0], o.sum((0, 2))) <= 1e-4 and _rel(sums[:, 1], (o * o).sum((0, 2))) <= 1e-4
    if ni * hw == 1:
        return  # torch refuses training-mode BatchNorm on one value per channel
    bn = torch.nn.BatchNorm2d(n).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    rm, rv = bn.running_mean.clone(), bn.running_var.clone()
    mi, sb = ops.bn_finalize(stats, ni * hw, bn.weight, bn.bias, rm, rv, 0.1, bn.eps)
    y = ops.bn_apply(out.view(ni, n, hw, 1), sb, relu=True)
    want = torch.relu(bn(out.float().view(ni, n, hw, 1)))
    assert _rel(y, want) <= 1e-2
    if ni * hw > 1:
        assert _rel(rm, bn.running_mean) <= 1e-4 and _rel(rv, bn.running_var) <= 1e-3
    assert _rel(mi[:, 0], o.mean((0, 2))) <= 1e-4
    # the packed transposed copy drives the input gradient
    g = torch.randn(ni, n, hw, device="cuda").to(BF)
    assert _rel(ops.pw_conv(g, w_kn), torch.matmul(w_nk.float().t(), g.float())) <= 1e-2


@pytest.mark.parametrize("shape", SHAPES)
def test_pw_conv_bn_relu_producer(shape):
    ni, k, n, hw = shape
    torch.manual_seed(1)
    x = torch.randn(ni, k, hw, device="cuda").to(BF)
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    sb = torch.stack([torch.rand(k, device="cuda") + 0.5, torch.randn(k, device="cuda")], dim=1).contiguous()
    a = torch.relu(x.float() * sb[:, 0].view(1, k, 1) + sb[:, 1].view(1, k, 1)).to(BF)
    ref = torch.matmul(w.to(BF).float(), a.float())
    assert _rel(ops.pw_conv(x, w, in_scale_bias=sb), ref) <= 1e-2
    g = torch.randn(ni, n, hw, device="cuda").to(BF)
    dw = ops.pw_conv_wgrad(g, x, in_scale_bias=sb)
    assert _rel(dw, torch.einsum("inp,ikp->nk", g.float(), a.float())) <= 1e-3


@pytest.mark.parametrize("shape", SHAPES)
def test_pw_conv_wgrad(shape):
    ni, k, n, hw = shape
    torch.manual_seed(2)
    x = torch.randn(ni, k, hw, device="cuda").to(BF)
    g = torch.randn(ni, n, hw, device="cuda").to(BF)
    dw = ops.pw_conv_wgrad(g, x)
    assert dw.shape == (n, k) and dw.dtype == torch.float32
    assert _rel(dw, torch.einsum("inp,ikp->nk", g.float(), x.float())) <= 1e-4
    dw2 = ops.pw_conv_wgrad(g, x)
    assert torch.equal(dw, dw2), "weight gradient must be deterministic (fixed-order reduction)"


# (clips, T, C, H, W, Cout)
SHIFT_SHAPES = [(2, 8, 72, 56, 56, 72), (2, 8, 144, 28, 28, 144), (2, 8, 288, 14, 14, 288), (2, 8, 576, 7, 7, 576),
                (1, 8, 54, 28, 28, 54), (2, 3, 10, 9, 11, 20), (1, 1, 16, 14, 14, 16), (1, 8, 24, 5, 3, 24)]


@pytest.mark.parametrize("shape", SHIFT_SHAPES)
@pytest.mark.parametrize("shift_kind", ["uniform1", "uniform3", "integer", "large"])
def test_shift3d_pw_conv(shape, shift_kind):
    n, t, c, h, w, cout = shape
    torch.manual_seed(3)
    x = torch.randn(n * t, c, h, w, device="cuda").to(BF)
    if shift_kind == "uniform1":
        shift = torch.rand(3, c, device="cuda") * 2 - 1
    elif shift_kind == "uniform3":
        shift = torch.rand(3, c, device="cuda") * 6 - 3
    elif shift_kind == "integer":
        shift = torch.randint(-2, 3, (3, c), device="cuda").float()
        shift[:, ::3] += 0.5
    else:
        shift = torch.rand(3, c, device="cuda") * 30 - 15
    wgt = torch.randn(cout, c, device="cuda") / c ** 0.5
    res = torch.randn(n * t, cout, h, w, device="cuda").to(BF)
    # the stand-alone sm_100a shift kernel (itself parity-tested against the oracle / reference goldens)
    shifted = rubiks_shift_3d_forward(x.view(n, t, c, h, w), shift, (1, 1, 1), 0).view(n * t, c, h, w)
    ref = ops.pw_conv(shifted, wgt, residual=res)
    out = ops.shift3d_pw_conv(x, shift, wgt, res, t)
    # same arithmetic (fp32 trilinear interpolation in the reference's association order, one rounding to bf16, then
    # the GEMM); only the compiler's FMA contraction inside the interpolation may differ between the two kernels
    assert _rel(out, ref) <= 4e-3, "fused shift+conv3 differs from shift kernel -> GEMM kernel"
    assert (out == ref).float().mean().item() >= 0.99
    assert _rel(out, torch.matmul(wgt.to(BF).float(), shifted.float().view(n * t, c, -1)).view_as(out) + res.float()) <= 1e-2
    g = torch.randn(n * t, cout, h, w, device="cuda").to(BF)
    dw = ops.shift3d_pw_conv_wgrad(g, x, shift, t)
    # `shifted` comes from the other kernel: an occasional 1-ulp bf16 difference in the recomputed operand is allowed
    err = _rel(dw, torch.einsum("inp,ikp->nk", g.float().flatten(2), shifted.float().flatten(2)))
    assert err <= 1e-3, err


def test_errors_are_reported():
    L = _lib.lib()
    x = torch.zeros(1, 8, 4, device="cuda", dtype=BF)
    w = torch.zeros(8, 8, device="cuda")
    out = torch.zeros(1, 8, 4, device="cuda", dtype=BF)
    rc = L.rb_pw_conv_forward(_lib.ptr(x), _lib.ptr(w), _lib.RB_F32, 0, None, _lib.ptr(out), _lib.RB_F32, 1, 8, 8, 4, None, None)
    assert rc == 2 and b"bf16" in L.rb_last_error()
    rc = L.rb_pw_conv_forward(_lib.ptr(x), _lib.ptr(w), _lib.RB_F16, 0, None, _lib.ptr(out), _lib.RB_BF16, 1, 8, 8, 4, None, None)
    assert rc == 1
    rc = L.rb_pw_conv_forward(None, _lib.ptr(w), _lib.RB_F32, 0, None, _lib.ptr(out), _lib.RB_BF16, 1, 8, 8, 4, None, None)
    assert rc == 1
    # empty batch is a no-op
    assert L.rb_pw_conv_forward(None, None, _lib.RB_F32, 0, None, None, _lib.RB_BF16, 0, 8, 8, 4, None, None) == 0
    dw = torch.ones(8, 8, device="cuda")
    assert L.rb_pw_conv_wgrad(None, None, _lib.ptr(dw), _lib.RB_BF16, 0, 8, 8, 4, None, None, 0, None) == 0
    torch.cuda.synchronize()
    assert float(dw.abs().sum()) == 0.0


def test_epilogue_bn_stats_block_chain():
    """Two chained identity blocks with fused.EPILOGUE_BN_STATS: the second block takes its bn1 statistics from the first
    block's conv3 epilogue; outputs, gradients and running statistics match the default schedule."""
    from rubiksnet_b200 import fused
    torch.manual_seed(11)
    net = rb.RubiksNet(tier="tiny", num_classes=5, num_frames=8).cuda().train()
    blocks = torch.nn.Sequential(net.backbone.layer3[1], net.backbone.layer3[2])
    x0 = torch.randn(16, 216, 14, 14, device="cuda").to(BF)
    res = []
    for flag in (False, True):
        fused.EPILOGUE_BN_STATS = flag
        try:
            sd = {k: v.clone() for k, v in blocks.state_dict().items()}
            blocks.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_()
            out = blocks(x)
            assert hasattr(out, "_rb_bn_stats") == flag
            out.float().square().mean().backward()
            res.append((out.detach().float(), x.grad.float(), {n: p.grad.float().clone() for n, p in blocks.named_parameters()},
                        blocks[1].bn1.running_var.clone()))
            blocks.load_state_dict(sd)
        finally:
            fused.EPILOGUE_BN_STATS = False
    (o0, g0, p0, rv0), (o1, g1, p1, rv1) = res
    assert _rel(o1, o0) <= 1e-2 and _rel(g1, g0) <= 2e-2 and _rel(rv1, rv0) <= 1e-3
    for name in p0:
        if not name.endswith("shift"):
            assert _rel(p1[name], p0[name]) <= 2e-2, name


@pytest.mark.parametrize("fuse_shift", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_whole_block_function_equals_module_graph(training, fuse_shift):
    """Identity-shortcut blocks as ONE autograd Function (shift fused into conv3, bn1 folded into conv2) vs the
    per-op fused path, same bf16 inputs: outputs, input gradient and every parameter gradient."""
    torch.manual_seed(5)
    net = rb.RubiksNet(tier="tiny", num_classes=7, num_frames=8).cuda()
    block = net.backbone.layer2[1]
    block.train(training)
    with torch.no_grad():
        for bn in (block.bn1, block.bn2):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    x0 = torch.randn(16, block.conv2.in_channels, 28, 28, device="cuda").to(BF)
    g = torch.randn(16, block.conv3.out_channels, 28, 28, device="cuda").to(BF)
    results = []
    from rubiksnet_b200 import fused
    for flag in (True, False):
        backbone.FUSED_WHOLE_BLOCK = flag
        fused.FUSE_SHIFT_CONV3 = fuse_shift
        try:
            sd = {k: v.clone() for k, v in block.state_dict().items()}
            block.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_()
            launches0 = _lib.launch_count()
            out = block(x)
            out.backward(g)
            n_launch = _lib.launch_count() - launches0
            results.append((out.detach().float(), x.grad.float(), {n: p.grad.float().clone() for n, p in block.named_parameters()},
                            block.bn2.running_var.clone(), n_launch))
            block.load_state_dict(sd)
        finally:
            backbone.FUSED_WHOLE_BLOCK = True
            fused.FUSE_SHIFT_CONV3 = False
    (o1, gx1, gp1, rv1, n1), (o0, gx0, gp0, rv0, n0) = results
    assert n1 < n0, "the whole-block Function must launch fewer kernels than the per-op path"
    assert _rel(o1, o0) <= 1e-2 and _rel(gx1, gx0) <= 2e-2
    assert _rel(rv1, rv0) <= 1e-3
    for name in gp0:
        if name.endswith("shift"):
            d = (gp1[name] - gp0[name]).abs()
            assert d.mean().item() <= 5e-3 and d.max().item() <= 0.25, (name, d.mean().item(), d.max().item())
        else:
            assert _rel(gp1[name], gp0[name]) <= 2e-2, name


@pytest.mark.parametrize("stage,hw", [("layer1", 56), ("layer3", 28), ("layer4", 14)])
@pytest.mark.parametrize("training", [True, False])
def test_down_sampling_block_function_equals_module_graph(stage, hw, training):
    """Down-sampling blocks (stride-(1,2,2) shift, stride-2 shortcut conv) as ONE autograd Function -- bn1 folded into conv2's
    producer, the shortcut fed from the sub-sampled quarter, its input gradient added in place -- vs the per-op path on
    the same bf16 inputs: output, input gradient, every parameter gradient, running statistics."""
    torch.manual_seed(7)
    net = rb.RubiksNet(tier="tiny", num_classes=7, num_frames=8).cuda()
    block = getattr(net.backbone, stage)[0]
    assert not isinstance(block.shortcut, torch.nn.Identity)
    block.train(training)
    with torch.no_grad():
        for bn in (block.bn1, block.bn2):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    x0 = torch.randn(16, block.conv2.in_channels, hw, hw, device="cuda").to(BF)
    g = torch.randn(16, block.conv3.out_channels, hw // 2, hw // 2, device="cuda").to(BF)
    results = []
    for flag in (True, False):
        backbone.FUSED_WHOLE_BLOCK = flag
        try:
            sd = {k: v.clone() for k, v in block.state_dict().items()}
            block.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_()
            out = block(x)
            fn_name = type(out.grad_fn).__name__
            out.backward(g)
            results.append((out.detach().float(), x.grad.float(), {n: p.grad.float().clone() for n, p in block.named_parameters()},
                            (block.bn1.running_mean.clone(), block.bn2.running_var.clone()), fn_name))
            block.load_state_dict(sd)
        finally:
            backbone.FUSED_WHOLE_BLOCK = True
    (o1, gx1, gp1, rs1, n1), (o0, gx0, gp0, rs0, n0) = results
    assert o1.shape == (16, block.conv3.out_channels, hw // 2, hw // 2)
    assert "RubiksDownBlockFn" in n1 and "RubiksDownBlockFn" not in n0, (n1, n0)
    assert _rel(o1, o0) <= 1e-2 and _rel(gx1, gx0) <= 2e-2
    for a, b in zip(rs1, rs0):
        assert _rel(a, b) <= 1e-3
    for name in gp0:
        if name.endswith("shift"):
            d = (gp1[name] - gp0[name]).abs()
            assert d.mean().item() <= 5e-3 and d.max().item() <= 0.25, (name, d.mean().item(), d.max().item())
        else:
            assert _rel(gp1[name], gp0[name]) <= 2e-2, name


@pytest.mark.parametrize("stage,idx,hw", [("layer1", 1, 56), ("layer3", 2, 14), ("layer4", 1, 7)])
@pytest.mark.parametrize("training", [True, False])
def test_aq_block_function_equals_module_graph(stage, idx, hw, training):
    """Identity-shortcut blocks of the attention-quantized variant (AttentionShift -> conv2, 2D shift) as ONE autograd
    Function vs the per-op path on the same bf16 inputs: output, input gradient, every parameter gradient (incl. the
    attention weights through the tap softmax), running statistics."""
    torch.manual_seed(8)
    net = rb.RubiksNet(tier="tiny", num_classes=7, num_frames=8, variant="rubiks3d-aq").cuda()
    block = getattr(net.backbone, stage)[idx]
    assert isinstance(block.shortcut, torch.nn.Identity)
    block.train(training)
    with torch.no_grad():
        for bn in (block.bn1, block.bn2):
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.uniform_(-0.3, 0.3)
            bn.running_mean.uniform_(-0.2, 0.2)
            bn.running_var.uniform_(0.5, 1.5)
    cin = block.conv2[1].in_channels
    x0 = torch.randn(16, cin, hw, hw, device="cuda").to(BF)
    g = torch.randn(16, block.conv3.out_channels, hw, hw, device="cuda").to(BF)
    results = []
    for flag in (True, False):
        backbone.FUSED_WHOLE_BLOCK = flag
        try:
            sd = {k: v.clone() for k, v in block.state_dict().items()}
            block.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_()
            out = block(x)
            fn_name = type(out.grad_fn).__name__
            out.backward(g)
            results.append((out.detach().float(), x.grad.float(), {n: p.grad.float().clone() for n, p in block.named_parameters()
                                                                   if p.grad is not None},
                            (block.bn1.running_mean.clone(), block.bn2.running_var.clone()), fn_name))
            block.load_state_dict(sd)
        finally:
            backbone.FUSED_WHOLE_BLOCK = True
    (o1, gx1, gp1, rs1, n1), (o0, gx0, gp0, rs0, n0) = results
    assert "RubiksAQBlockFn" in n1 and "RubiksAQBlockFn" not in n0, (n1, n0)
    assert set(gp1) == set(gp0) and any(k.endswith("conv2.0.weight") for k in gp1)
    assert _rel(o1, o0) <= 1e-2 and _rel(gx1, gx0) <= 2e-2
    for a, b in zip(rs1, rs0):
        assert _rel(a, b) <= 1e-3
    for name in gp0:
        if name.endswith("shift"):
            d = (gp1[name] - gp0[name]).abs()
            assert d.mean().item() <= 5e-3 and d.max().item() <= 0.25, (name, d.mean().item(), d.max().item())
        else:
            assert _rel(gp1[name], gp0[name]) <= 2e-2, name


@pytest.mark.parametrize("fuse_shift", [False, True])
def test_training_step_uses_tensor_core_blocks(fuse_shift):
    """A bf16 autocast training step of RubiksNet-Tiny routes its identity-shortcut blocks through the tcgen05 kernels
    (and, with FUSE_SHIFT_CONV3, through the single shift+conv3 launch)."""
    from rubiksnet_b200 import fused
    torch.manual_seed(6)
    net = rb.RubiksNet(tier="tiny", num_classes=5, num_frames=8).cuda().train()
    clips = torch.randn(1, 8, 3, 224, 224, device="cuda")
    fused.FUSE_SHIFT_CONV3 = fuse_shift
    try:
        _lib.timing.start()
        with torch.autocast("cuda", dtype=BF):
            loss = net(clips).float().square().mean()
        loss.backward()
        agg = _lib.timing.stop()
    finally:
        fused.FUSE_SHIFT_CONV3 = False
    # 13 identity-shortcut blocks + 4 down-sampling blocks, all as whole-block Functions with bn1 folded into conv2's producer
    assert agg["pw_conv<bn+relu>"]["launches"] == 17
    if fuse_shift:  # the single shift+conv3 launch exists for the stride-1 (identity) blocks
        assert agg["shift3d_pw_conv"]["launches"] == 13 and agg["shift3d_pw_conv_wgrad"]["launches"] == 13
        assert agg["pw_conv<+residual>"]["launches"] == 4
    else:
        assert agg["pw_conv<+residual>"]["launches"] == 17 and "shift3d_pw_conv" not in agg
    assert torch.isfinite(loss).item()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
