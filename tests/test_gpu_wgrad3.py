"""Tensor-map TMA schedule of the 1x1-conv weight gradient (csrc/pw_wgrad3.cu) against the first tcgen05 kernel on the same
bf16 tensors (same arithmetic: bf16 operands, fp32 accumulation in tensor memory, fixed-order split reduction) and against
fp32 math.  Maps whose row pitch is not a multiple of 16 bytes (14x14, 7x7, ...) must keep taking the first kernel."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402,F401

from rubiksnet_b200 import _lib, ops  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]
BF = torch.bfloat16

GEOMS = [  # (NI, K = input channels, N = output channels, H, W)
    (8, 72, 72, 112, 112),   # layer0
    (8, 72, 72, 56, 56),     # layer1.x
    (8, 144, 144, 28, 28),   # layer2.x
    (16, 288, 288, 16, 16),  # three 128-row M tiles x two column blocks, 512 TMEM columns
    (16, 576, 576, 8, 8),    # 2 x 4 blocks, one chunk per image
    (16, 288, 288, 14, 14),  # layer3.x: 392-byte rows -> first kernel
    (16, 576, 576, 7, 7),    # layer4.x: 98-byte rows -> first kernel
    (8, 72, 144, 56, 56),    # down-sampling conv2
    (8, 72, 144, 28, 28),    # shortcut conv of a down-sampling block on the sub-sampled map
    (8, 288, 576, 14, 14),
    (4, 32, 72, 112, 112),   # conv1 on the patch matrix
    (3, 32, 64, 6, 5),       # HW = 30 -> first kernel
    (3, 32, 64, 6, 4),       # HW = 24: a single partly filled chunk per image
    (5, 16, 8, 16, 8),       # smallest operands, HW a multiple of the chunk
    (2, 256, 128, 16, 16),
    (3, 48, 80, 10, 10),     # HW = 100 -> first kernel
    (3, 48, 80, 13, 8),      # HW = 104: 40 valid pixels in the last chunk
    (2, 520, 264, 8, 8),     # boxes of 260 > 256 rows are split in two
]


def _run(g, x, sb, tma):
    _lib.set_pw_tma(tma)
    try:
        dw = ops.pw_conv_wgrad(g, x, in_scale_bias=sb)
        torch.cuda.synchronize()
        return dw
    finally:
        _lib.set_pw_tma(True)


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("bn", [False, True])
def test_wgrad_tma_matches_first_kernel(geom, bn):
    ni, k, n, h, w_ = geom
    torch.manual_seed(sum(geom) + bn)
    x = torch.randn(ni, k, h, w_, device="cuda").to(BF)
    g = torch.randn(ni, n, h, w_, device="cuda").to(BF)
    sb = torch.stack([torch.rand(k, device="cuda") + 0.5, torch.randn(k, device="cuda")], dim=1).contiguous() if bn else None
    got = _run(g, x, sb, True)
    want = _run(g, x, sb, False)
    a = x.float()
    if bn:
        a = torch.relu(a * sb[:, 0].view(1, -1, 1, 1) + sb[:, 1].view(1, -1, 1, 1)).to(BF).float()
    ref = torch.einsum("inhw,ikhw->nk", g.float(), a)
    scale = max(1.0, ref.abs().max().item())
    assert got.shape == (n, k)
    assert (got - ref).abs().max().item() <= 2e-3 * scale, ((got - ref).abs().max().item(), scale)
    assert (got - want).abs().max().item() <= 2e-3 * scale


def test_wgrad_tma_is_taken_and_deterministic():
    torch.manual_seed(3)
    x = torch.randn(16, 144, 28, 28, device="cuda").to(BF)
    g = torch.randn(16, 144, 28, 28, device="cuda").to(BF)
    a = ops.pw_conv_wgrad(g, x)
    b = ops.pw_conv_wgrad(g, x)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # operands that are not 16-byte aligned take the first kernel and agree
    buf = torch.randn(16 * 144 * 784 + 4, device="cuda").to(BF)
    xs = buf[4:].view(16, 144, 28, 28)
    xs.copy_(x)
    assert xs.data_ptr() % 16 == 8
    c = ops.pw_conv_wgrad(g, xs)
    torch.cuda.synchronize()
    assert (c - a).abs().max().item() <= 2e-3 * max(1.0, a.abs().max().item())
    # schedule switch off: first kernel, same result up to the summation order of the pixel splits
    _lib.set_pw_tma(False)
    try:
        d = ops.pw_conv_wgrad(g, x)
    finally:
        _lib.set_pw_tma(True)
    torch.cuda.synchronize()
    assert (d - a).abs().max().item() <= 2e-3 * max(1.0, a.abs().max().item())
