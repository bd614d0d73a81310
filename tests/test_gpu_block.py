"""GPU tests of the RubiksShiftBlock path: fused BN+ReLU kernels and batched-GEMM 1x1 convs against
torch.nn modules, the fused block against the plain module graph, and the whole network against the
reference package + extension (baseline/_ref) on a pretrained checkpoint."""
import contextlib
import io
import os
import sys

import pytest
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200 import backbone, fused  # noqa: E402

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / max(1.0, b.double().abs().max().item())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 1e-2), (torch.float16, 1e-2)])
@pytest.mark.parametrize("shape", [(16, 72, 28, 28), (8, 288, 14, 14), (8, 54, 7, 7), (3, 5, 3, 3), (4, 9, 56, 57)])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_act_training(shape, dtype, tol, relu):
    torch.manual_seed(0)
    c = shape[1]
    x = (torch.randn(shape, device="cuda") * 1.5 + 0.3).to(dtype)
    bn_ref = nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn_ref.weight.uniform_(0.5, 1.5)
        bn_ref.bias.uniform_(-0.5, 0.5)
    bn_new = nn.BatchNorm2d(c).cuda()
    bn_new.load_state_dict(bn_ref.state_dict())
    xr = x.float().clone().requires_grad_()
    yr = bn_ref(xr)
    yr = torch.relu(yr) if relu else yr
    g = torch.randn_like(yr)
    yr.backward(g)
    xn = x.detach().clone().requires_grad_()
    yn = fused.bn_act(xn, bn_new, relu=relu)
    yn.backward(g.to(dtype))
    assert yn.dtype == dtype
    assert _rel(yn, yr) <= tol
    assert _rel(xn.grad, xr.grad) <= tol
    assert _rel(bn_new.weight.grad, bn_ref.weight.grad) <= max(tol, 2e-3 if dtype != torch.float32 else tol)
    assert _rel(bn_new.bias.grad, bn_ref.bias.grad) <= max(tol, 2e-3 if dtype != torch.float32 else tol)
    assert _rel(bn_new.running_mean, bn_ref.running_mean) <= 1e-4
    assert _rel(bn_new.running_var, bn_ref.running_var) <= 1e-4
    assert int(bn_new.num_batches_tracked) == 1


def test_bn_act_eval_mode():
    torch.manual_seed(1)
    bn = nn.BatchNorm2d(12).cuda()
    with torch.no_grad():
        bn.running_mean.uniform_(-1, 1)
        bn.running_var.uniform_(0.5, 2)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
    bn.eval()
    x = torch.randn(4, 12, 9, 9, device="cuda", requires_grad=True)
    xr = x.detach().clone().requires_grad_()
    y = fused.bn_act(x, bn, relu=True)
    yr = torch.relu(bn(xr))
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g)
    assert _rel(y, yr) <= 1e-5 and _rel(x.grad, xr.grad) <= 1e-5
    assert int(bn.num_batches_tracked) == 0


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-3), (torch.bfloat16, 2e-2)])  # fp32 matmul may use TF32
def test_conv1x1_matches_conv2d(dtype, tol):
    torch.manual_seed(2)
    x = torch.randn(8, 72, 14, 14, device="cuda").to(dtype).requires_grad_()
    w = (torch.randn(144, 72, 1, 1, device="cuda") * 0.1).requires_grad_()
    res = torch.randn(8, 144, 7, 7, device="cuda").to(dtype).requires_grad_()
    out = fused.conv1x1(x, w, residual=res, stride=2)
    ref = torch.nn.functional.conv2d(x.float(), w, stride=2) + res.float()
    g = torch.randn_like(ref)
    gx, gw, gr = torch.autograd.grad(out, (x, w, res), g.to(dtype))
    rx, rw, rr = torch.autograd.grad(ref, (x, w, res), g)
    assert _rel(out, ref) <= tol and _rel(gx, rx) <= tol and _rel(gw, rw) <= tol and _rel(gr, rr) <= tol
    assert gw.dtype == torch.float32


@pytest.mark.parametrize("variant", ["rubiks3d", "rubiks3d-aq"])
def test_fused_block_equals_module_graph(variant):
    torch.manual_seed(3)
    net = rb.RubiksNet(tier="tiny", num_classes=7, num_frames=8, variant=variant).cuda().train()
    clips = torch.randn(1, 8, 3, 224, 224, device="cuda")
    results = []
    for flag in (True, False):
        backbone.FUSED_BLOCK = flag
        try:
            sd = {k: v.clone() for k, v in net.state_dict().items()}
            net.zero_grad(set_to_none=True)
            logits = net(clips)
            logits.square().sum().backward()
            grads = {n: p.grad.clone() for n, p in net.named_parameters() if p.requires_grad}
            results.append((logits.detach().clone(), grads, net.backbone.bn_last.running_var.clone()))
            net.load_state_dict(sd)  # undo the running-stat update
        finally:
            backbone.FUSED_BLOCK = True
    (l1, g1, rv1), (l0, g0, rv0) = results
    assert _rel(l1, l0) <= 5e-3
    assert _rel(rv1, rv0) <= 1e-3
    for name in ("backbone.conv1.weight", "backbone.layer2.1.conv2.weight" if variant == "rubiks3d" else "backbone.layer2.1.conv2.1.weight",
                 "backbone.layer3.0.bn2.weight", "new_fc.weight"):
        assert _rel(g1[name], g0[name]) <= 2e-2, name


def _reference_package():
    ref = os.path.join(REPO, "baseline", "_ref")
    ckpt = os.path.join(ref, "pretrained", "ssv2_tiny.pth.tar")
    if not (os.path.isdir(os.path.join(ref, "rubiksnet")) and os.path.exists(ckpt)):
        return None, None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from rubiksnet.models import RubiksNet as RefNet
        return RefNet, ckpt
    except Exception:
        return None, None


def test_whole_model_matches_reference_on_pretrained_checkpoint():
    """Logits (eval) and a training step's gradients of RubiksNet-Tiny (ssv2 checkpoint): this package vs the
    reference package running its own CUDA extension."""
    RefNet, ckpt = _reference_package()
    if RefNet is None:
        pytest.skip("baseline/_ref (reference package + checkpoint) not available on this box")
    torch.manual_seed(4)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefNet.load_pretrained(ckpt).cuda()
    new = rb.RubiksNet.load_pretrained(ckpt).cuda()
    clips = torch.randn(2, 8, 3, 224, 224, device="cuda")
    old_tf32 = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ref.eval()
        new.eval()
        with torch.no_grad():
            lr, ln = ref(clips), new(clips)
        assert _rel(ln, lr) <= 1e-3, "eval logits differ from the reference"
        assert torch.equal(ln.argmax(1), lr.argmax(1))
        ref.train()
        new.train()
        labels = torch.tensor([3, 100], device="cuda")
        torch.nn.functional.cross_entropy(ref(clips), labels).backward()
        torch.nn.functional.cross_entropy(new(clips), labels).backward()
        gr = dict(ref.named_parameters())
        for name, p in new.named_parameters():
            if name in ("new_fc.weight", "backbone.conv1.weight", "backbone.layer3.2.conv3.weight"):
                assert _rel(p.grad, gr[name].grad) <= 5e-3, name
            if name.endswith("shift"):
                # per-channel L2-normalised gradients: channels whose raw gradient is tiny amplify rounding noise
                # (the reference itself accumulates with fp32 atomics in arbitrary order)
                d = (p.grad - gr[name].grad).abs()
                assert d.mean().item() <= 2e-3 and d.max().item() <= 0.2, (name, d.mean().item(), d.max().item())
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("shape", [(16, 72, 14, 14), (8, 144, 7, 7), (4, 24, 28, 28), (3, 36, 5, 3)])
def test_se_gate_matches_module(shape, dtype, tol):
    """fused.se_gate (librubiks_b200 pool / rescale kernels) vs the SELayer module graph: output, input gradient and both
    Linear weight gradients."""
    torch.manual_seed(8)
    se = backbone.SELayer(shape[1], reduction=12).cuda()
    x = torch.randn(shape, device="cuda").to(dtype)
    g = torch.randn(shape, device="cuda").to(dtype)
    xr = x.float().clone().requires_grad_()
    yr = se(xr)
    yr.backward(g.float())
    ref_w = [p.grad.clone() for p in se.parameters()]
    se.zero_grad(set_to_none=True)
    xn = x.clone().requires_grad_()
    yn = fused.se_gate(xn, se)
    yn.backward(g)
    assert yn.dtype == dtype
    assert _rel(yn, yr) <= tol and _rel(xn.grad, xr.grad) <= tol
    for p, r in zip(se.parameters(), ref_w):
        assert _rel(p.grad, r) <= max(tol, 1e-4)


def test_se_gate_under_autocast():
    """Inside a bf16 autocast region the gate MLP must stay fp32: the rescale kernel reads float gates."""
    torch.manual_seed(10)
    se = backbone.SELayer(72, reduction=12).cuda()
    x = torch.randn(16, 72, 14, 14, device="cuda").bfloat16()
    g = torch.randn_like(x)
    xr = x.float().clone().requires_grad_()
    se(xr).backward(g.float())
    ref_w = [p.grad.clone() for p in se.parameters()]
    se.zero_grad(set_to_none=True)
    xn = x.clone().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yn = fused.se_gate(xn, se)
    yn.backward(g)
    assert _rel(yn, se(x.float())) <= 1e-2 and _rel(xn.grad, xr.grad) <= 1e-2
    for p, r in zip(se.parameters(), ref_w):
        assert _rel(p.grad, r) <= 1e-2


def test_small_tier_uses_se_kernels():
    from rubiksnet_b200 import _lib
    torch.manual_seed(9)
    net = rb.RubiksNet(tier="small", num_classes=5, num_frames=8).cuda().train()
    clips = torch.randn(1, 8, 3, 224, 224, device="cuda")
    _lib.timing.start()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = net(clips).float().square().mean()
    loss.backward()
    agg = _lib.timing.stop()
    assert agg["se_plane_scale"]["launches"] == 2 * 17 and agg["se_plane_reduce"]["launches"] == 2 * 17
    assert torch.isfinite(loss).item() and all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("cout,hw,stride", [(72, 224, 2), (54, 64, 2), (24, 32, 1)])
def test_stem_conv_bf16_autocast_matches_conv2d(cout, hw, stride):
    """conv1 as im2col + tcgen05 GEMM (+ tcgen05 weight gradient) vs nn.Conv2d in fp32 on the same clips."""
    torch.manual_seed(11)
    conv = nn.Conv2d(3, cout, 3, stride=stride, padding=1, bias=False).cuda()
    x = torch.randn(6, 3, hw, hw, device="cuda")
    want = conv(x)
    g = torch.randn_like(want)
    want.backward(g)
    gw_ref = conv.weight.grad.clone()
    conv.weight.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = fused.stem_conv(conv, x)
    assert got.dtype == torch.bfloat16 and got.shape == want.shape
    got.backward(g.bfloat16())
    assert _rel(got, want) <= 1e-2
    assert _rel(conv.weight.grad, gw_ref) <= 1e-2


def test_stem_conv_fp32_inference_and_fallbacks():
    torch.manual_seed(12)
    conv = nn.Conv2d(3, 54, 3, stride=2, padding=1, bias=False).cuda()
    x = torch.randn(5, 3, 224, 224, device="cuda")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad():
            got = fused.stem_conv(conv, x)  # kind::tf32 GEMM on the fp32 patch matrix
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            want = conv(x)
            assert torch.equal(fused.stem_conv(conv, x), want)  # TF32 off: the module itself
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    assert got.dtype == torch.float32 and _rel(got, want) <= 2e-3
    xg = x.clone().requires_grad_()  # an input that needs a gradient: the module itself
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = fused.stem_conv(conv, xg)
    y.float().sum().backward()
    assert xg.grad is not None


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(256, 288, 14, 14), (64, 576, 7, 7), (24, 144, 14, 14), (40, 130, 5, 6), (8, 128, 28, 28)])
@pytest.mark.parametrize("residual", [False, True])
def test_bn_channel_resident_passes_match_streaming(shape, dtype, residual):
    """Small maps run the BatchNorm forward as ONE channel-resident launch (csrc/bn_relu.cu: k_bn_fwd_resident); same
    statistics, outputs and -- through the saved statistics -- gradients as the streaming reduce / finalize / apply passes."""
    from rubiksnet_b200 import _lib, ops
    torch.manual_seed(13)
    c = shape[1]
    x = (torch.randn(shape, device="cuda") * 1.3 + 0.4).to(dtype)
    dy = torch.randn(shape, device="cuda").to(dtype)
    res = torch.randn(shape, device="cuda").to(dtype) if residual else None
    g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.3
    out = {}
    for mode in (True, False):
        _lib.set_bn_resident(mode)
        try:
            rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
            n0 = _lib.launch_count()
            y, mi, sb = ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=True, apply=True)
            nf = _lib.launch_count() - n0
            dx, dg, db = ops.bn_backward(x, dy, res, g, mi, sb, True, relu=True)
            nb = _lib.launch_count() - n0 - nf
            out[mode] = (y.float(), mi, sb, rm, rv, dx.float(), dg, db, nf, nb)
        finally:
            _lib.set_bn_resident(True)
    a, s = out[True], out[False]
    assert a[8] == 1 and s[8] == 3 and a[9] == s[9] == 3, (a[8:], s[8:])  # forward: one launch instead of three
    for i, tol in ((0, 1e-2), (1, 1e-5), (2, 1e-5), (3, 1e-6), (4, 1e-6), (5, 1e-2), (6, 2e-4), (7, 2e-4)):
        assert _rel(a[i], s[i]) <= tol, (i, _rel(a[i], s[i]))
    # and against torch in fp32
    bn = nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        bn.weight.copy_(g)
        bn.bias.copy_(b)
    xr = x.float().requires_grad_()
    yr = torch.relu(bn(xr))
    yr.backward(dy.float())
    assert _rel(a[0], yr) <= 1e-2
    assert _rel(a[5], xr.grad + (res.float() if residual else 0)) <= 1e-2
    assert _rel(a[3], bn.running_mean) <= 1e-4 and _rel(a[4], bn.running_var) <= 1e-4


@pytest.mark.parametrize("variant", ["rubiks3d", "rubiks3d-aq"])
def test_every_batchnorm_counter_advances_once_per_training_forward(variant):
    """The backbone bumps all num_batches_tracked counters with one multi-tensor add; the per-layer code must then skip them
    (and still count on its own when a block is called outside a backbone forward, or in eval mode not at all)."""
    torch.manual_seed(0)
    net = rb.RubiksNet(tier="tiny", num_classes=5, num_frames=8, variant=variant).cuda().train()
    x = torch.randn(8, 3, 224, 224, device="cuda")
    bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        net(x)
        net(x)
    assert all(int(m.num_batches_tracked) == 2 for m in bns), sorted({int(m.num_batches_tracked) for m in bns})
    net.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        net(x)
    assert all(int(m.num_batches_tracked) == 2 for m in bns)
    blk = net.backbone.layer0[0]
    blk.train()
    xb = torch.randn(8, blk.bn1.num_features, 16, 16, device="cuda").to(torch.bfloat16)
    blk(xb)
    assert int(blk.bn1.num_batches_tracked) == 3 and int(blk.bn2.num_batches_tracked) == 3
