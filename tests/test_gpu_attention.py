"""GPU parity tests for the AttentionShift kernels (C ABI) against the numpy oracle and the reference
module's recorded outputs."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
from helpers import assert_close  # noqa: E402

import rubiksnet_b200 as rb  # noqa: E402
from rubiksnet_b200.attention_shift import attention_shift_mix  # noqa: E402

pytestmark = pytest.mark.gpu


def test_attention_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention_golden.npz"))
    for name in ("small", "sharp", "t1"):
        n, t, temp = g[name + "/meta"]
        m = rb.AttentionShift(int(t), channels=g[name + "/weight"].shape[0]).cuda().double()
        m.weight.data.copy_(torch.from_numpy(g[name + "/weight"]))
        m.T.data.fill_(float(temp))
        x = torch.from_numpy(g[name + "/x"]).cuda().requires_grad_()
        out = m(x)
        out.backward(torch.from_numpy(g[name + "/og"]).cuda())
        assert_close(out.detach().cpu().numpy(), g[name + "/out"], 1e-6, name + " out")   # taps are fp32
        assert_close(x.grad.cpu().numpy(), g[name + "/gx"], 1e-6, name + " gx")
        assert_close(m.weight.grad.cpu().numpy(), g[name + "/gweight"], 1e-5, name + " gweight")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2), (torch.float16, 1e-2)])
@pytest.mark.parametrize("shape", [(2, 8, 72, 28, 28), (3, 8, 13, 7, 7), (1, 4, 5, 57, 3)])
def test_attention_mix_vs_oracle(shape, dtype, tol):
    n, t, c, h, w = shape
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((n * t, c, h, w)).astype(np.float32)).cuda().to(dtype).requires_grad_()
    weight = rng.random((c, 3)).astype(np.float32)
    taps = torch.from_numpy(oracle.attention_taps(weight, 2.0).astype(np.float32)).cuda().requires_grad_()
    og = torch.from_numpy(rng.standard_normal((n * t, c, h, w)).astype(np.float32)).cuda().to(dtype)
    out = attention_shift_mix(x, taps, t)
    out.backward(og)
    xr, gr = x.detach().double().cpu().numpy(), og.double().cpu().numpy()
    a = taps.detach().double().cpu().numpy()
    xv, gv = xr.reshape(n, t, c, h, w), gr.reshape(n, t, c, h, w)
    exp = a[None, None, :, 1, None, None] * xv
    exp[:, 1:] += a[None, None, :, 0, None, None] * xv[:, :-1]
    exp[:, :-1] += a[None, None, :, 2, None, None] * xv[:, 1:]
    assert_close(out.detach().double().cpu().numpy(), exp.reshape(xr.shape), tol, "out")
    ga = np.stack([(gv[:, 1:] * xv[:, :-1]).sum((0, 1, 3, 4)), (gv * xv).sum((0, 1, 3, 4)),
                   (gv[:, :-1] * xv[:, 1:]).sum((0, 1, 3, 4))], axis=1)
    assert_close(taps.grad.double().cpu().numpy(), ga, 1e-4, "taps grad")
    gx = a[None, None, :, 1, None, None] * gv
    gx[:, :-1] += a[None, None, :, 0, None, None] * gv[:, 1:]
    gx[:, 1:] += a[None, None, :, 2, None, None] * gv[:, :-1]
    assert_close(x.grad.double().cpu().numpy(), gx.reshape(xr.shape), tol, "x grad")


@pytest.mark.parametrize("shape", [(16, 288, 14, 14), (16, 72, 28, 28), (8, 40, 7, 7), (24, 6, 5, 3)])
def test_bn_folded_mix_is_bit_identical_to_the_two_pass_path(shape):
    """rb_bn_attention_shift_{forward,backward}: relu(bn1(x)) applied as the kernels load == mixing the tensor that
    rb_bn_act_forward would have written (same FMA, same rounding), for the output, the input gradient and the tap gradient."""
    from rubiksnet_b200 import ops
    from rubiksnet_b200.attention_shift import attention_mix_backward, attention_mix_forward
    torch.manual_seed(sum(shape))
    nt, c, h, w = shape
    frames = 8
    x = torch.randn(nt, c, h, w, device="cuda").to(torch.bfloat16)
    gamma = torch.rand(c, device="cuda") + 0.5
    beta = torch.randn(c, device="cuda") * 0.3
    taps = torch.softmax(torch.randn(c, 3, device="cuda"), dim=1).contiguous()
    o, _, sb = ops.bn_forward(x, gamma, beta, None, None, True, 0.1, 1e-5, relu=True, apply=True)
    want = attention_mix_forward(o, taps, frames)
    got = attention_mix_forward(x, taps, frames, in_scale_bias=sb)
    assert torch.equal(got, want)
    g = torch.randn_like(x)
    gx_w, gt_w = attention_mix_backward(o, taps, g, frames)
    gx_g, gt_g = attention_mix_backward(x, taps, g, frames, in_scale_bias=sb)
    torch.cuda.synchronize()
    assert torch.equal(gx_g, gx_w)
    assert torch.equal(gt_g, gt_w)
