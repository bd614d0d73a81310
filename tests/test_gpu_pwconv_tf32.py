"""tcgen05 kind::tf32 pointwise convolution (csrc/pw_conv_tf32.cu, rb_pw_conv_forward_f32) against fp64 math on the
same fp32 inputs: plain, with the bn1+relu producer, with the bn2+relu epilogue and with the shortcut added -- the three
launches of an eval-mode RubiksShiftBlock (rubiksnet/backbone.py:123-135) -- and the whole eval-mode network against
the nn.Module graph.  Tolerance: TF32 keeps a 10-bit mantissa (2^-11 relative per operand), as do the cuDNN
convolutions the reference runs by default; 2e-3 of the output scale, written below."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402,F401

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200 import backbone, ops  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 2e-3


def _ref(x, w, res, in_sb, out_sb, relu):
    a = x.double()
    if in_sb is not None:
        a = torch.relu(a * in_sb[:, 0].double().view(1, -1, 1, 1) + in_sb[:, 1].double().view(1, -1, 1, 1))
    y = torch.einsum("nk,ikhw->inhw", w.double(), a)
    if out_sb is not None:
        y = y * out_sb[:, 0].double().view(1, -1, 1, 1) + out_sb[:, 1].double().view(1, -1, 1, 1)
    if relu:
        y = torch.relu(y)
    if res is not None:
        y = y + res.double()
    return y


GEOMS = [  # (NI, K, N, H, W)
    (4, 54, 54, 112, 112),   # tiny layer0
    (8, 54, 54, 56, 56),     # tiny layer1.x
    (8, 54, 108, 56, 56),    # layer2.0 conv2
    (16, 108, 108, 28, 28),
    (16, 216, 216, 14, 14),
    (16, 216, 432, 14, 14),  # layer4.0 conv2: four output-channel slices
    (16, 432, 432, 7, 7),    # contraction split over two launches, 49-pixel planes (scalar accesses)
    (3, 5, 3, 3, 3),         # a single partial tile
    (2, 72, 72, 14, 15),     # HW % 4 != 0
    (5, 304, 40, 6, 6),      # longest single-launch contraction
    (5, 320, 40, 6, 6),      # just beyond it
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("mode", ["plain", "in_bn", "out_bn_relu", "res", "all"])
def test_pw_conv_f32_matches_fp64(geom, mode):
    ni, k, n, h, w_ = geom
    torch.manual_seed(sum(geom) + len(mode))
    x = torch.randn(ni, k, h, w_, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    res = torch.randn(ni, n, h, w_, device="cuda") if mode in ("res", "all") else None
    in_sb = (torch.stack([torch.rand(k, device="cuda") + 0.5, torch.randn(k, device="cuda") * 0.5], dim=1).contiguous()
             if mode in ("in_bn", "all") else None)
    out_sb = (torch.stack([torch.rand(n, device="cuda") + 0.5, torch.randn(n, device="cuda") * 0.5], dim=1).contiguous()
              if mode in ("out_bn_relu", "all") else None)
    relu = mode in ("out_bn_relu", "all")
    got = ops.pw_conv_f32(x, w, residual=res, in_scale_bias=in_sb, out_scale_bias=out_sb, relu=relu)
    want = _ref(x, w, res, in_sb, out_sb, relu)
    assert got.dtype == torch.float32 and got.shape == want.shape
    err = (got.double() - want).abs().max().item()
    assert err <= TOL * max(1.0, want.abs().max().item()), (err, want.abs().max().item())
    # resident-weight flag: same arithmetic, weights staged before the dependency wait
    got2 = ops.pw_conv_f32(x, w, residual=res, in_scale_bias=in_sb, out_scale_bias=out_sb, relu=relu, resident=True)
    assert torch.equal(got, got2)


def test_pw_conv_f32_residual_in_place_and_unaligned():
    torch.manual_seed(3)
    base = torch.randn(4 * 72 * 196 + 1, device="cuda")
    x = base[1:].view(4, 72, 14, 14)  # 4-byte aligned only: scalar path
    w = torch.randn(72, 72, device="cuda") / 72 ** 0.5
    res = torch.randn(4, 72, 14, 14, device="cuda")
    got = ops.pw_conv_f32(x, w, residual=res)
    want = _ref(x, w, res, None, None, False)
    assert (got.double() - want).abs().max().item() <= TOL * want.abs().max().item()


@pytest.mark.parametrize("tier,variant", [("tiny", "rubiks3d"), ("tiny", "rubiks3d-aq"), ("small", "rubiks3d")])
def test_eval_network_fp32_matches_module_graph(tier, variant):
    """Eval-mode fp32 network on the fused tf32 path vs the plain nn.Module graph (cuDNN convs with TF32 off)."""
    torch.manual_seed(4)
    net = rb.RubiksNet(tier=tier, variant=variant, num_classes=11, num_frames=8).cuda()
    for m in net.modules():  # non-trivial running statistics / affine parameters
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    net.eval()
    clips = torch.randn(2, 8, 3, 224, 224, device="cuda")
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, backbone.FUSED_BLOCK)
    try:
        with torch.no_grad():
            got = net(clips)
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            backbone.FUSED_BLOCK = False
            want = net(clips)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, backbone.FUSED_BLOCK = old
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-2 * max(1.0, scale), ((got - want).abs().max().item(), scale)
