"""CPU tests (-m "not gpu") of the oracle: against the reference's own CUDA outputs recorded on a B200
(tests/golden/shift_golden.npz), against the reference's AttentionShift module, against an independent
autograd formulation, and hand-computed known answers for the exact-integer gradient rules."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
from golden.make_golden import cases2d, cases3d  # noqa: E402
from helpers import assert_close, load_golden, make_shift  # noqa: E402
from ref_torch import shift2d, shift3d  # noqa: E402

GOLD = load_golden("shift_golden.npz")
needs_golden = pytest.mark.skipif(GOLD is None, reason="tests/golden/shift_golden.npz not generated yet")


@needs_golden
@pytest.mark.parametrize("name", sorted(cases3d()))
def test_oracle_matches_reference_cuda_3d(name):
    cfg = cases3d()[name]
    g = lambda k: GOLD["3d/%s/%s" % (name, k)]
    tol = 1e-10 if cfg.get("dtype") == "float64" else 2e-5
    out = oracle.shift3d_forward(g("x"), g("shift"), cfg["stride"], cfg["padding"], cfg.get("quantize", False))
    assert_close(out, g("out"), tol, name + " out")
    gin, gs = oracle.shift3d_backward(g("x"), g("shift"), g("og"), cfg["stride"], cfg["padding"],
                                      cfg.get("normalize_grad", True), cfg.get("t_factor", 1.0),
                                      cfg.get("quantize", False))
    assert_close(gin, g("gin"), tol, name + " gin")
    assert_close(gs, g("gshift"), 1e-9 if cfg.get("dtype") == "float64" else 1e-4, name + " gshift")
    if cfg.get("quantize"):
        np.testing.assert_array_equal(out, g("out"))  # pure copies: bit exact
        np.testing.assert_array_equal(gin, g("gin"))


@needs_golden
@pytest.mark.parametrize("name", sorted(cases2d()))
def test_oracle_matches_reference_cuda_2d(name):
    cfg = cases2d()[name]
    g = lambda k: GOLD["2d/%s/%s" % (name, k)]
    tol = 1e-10 if cfg.get("dtype") == "float64" else 2e-5
    out = oracle.shift2d_forward(g("x"), g("shift"), cfg["stride"], cfg["padding"], cfg.get("quantize", False))
    assert_close(out, g("out"), tol, name + " out")
    gin, gs = oracle.shift2d_backward(g("x"), g("shift"), g("og"), cfg["stride"], cfg["padding"],
                                      cfg.get("normalize_grad", True), cfg.get("enable_shift_grad", True),
                                      cfg.get("quantize", False))
    assert_close(gin, g("gin"), tol, name + " gin")
    assert_close(gs, g("gshift"), 1e-9 if cfg.get("dtype") == "float64" else 1e-4, name + " gshift")


@pytest.mark.parametrize("stride,padding", [((1, 1, 1), (0, 0, 0)), ((1, 2, 2), (0, 0, 0)), ((2, 2, 1), (1, 1, 0)),
                                            ((1, 3, 2), (0, 2, 1))])
def test_oracle_3d_vs_autograd(stride, padding):
    torch.manual_seed(1)
    x = torch.randn(2, 5, 4, 7, 6, dtype=torch.float64, requires_grad=True)
    s = (torch.rand(3, 4, dtype=torch.float64) * 5 - 2.5).requires_grad_()
    out = shift3d(x, s, stride, padding)
    og = torch.randn_like(out)
    out.backward(og)
    assert_close(oracle.shift3d_forward(x, s, stride, padding), out.detach().numpy(), 1e-12, "out")
    gin, gs = oracle.shift3d_backward(x, s, og, stride, padding, normalize_grad=False)
    assert_close(gin, x.grad.numpy(), 1e-12, "gin")
    assert_close(gs, s.grad.numpy(), 1e-11, "gshift")


@pytest.mark.parametrize("stride,padding", [((1, 1), (0, 0)), ((2, 2), (0, 0)), ((2, 1), (1, 0))])
def test_oracle_2d_vs_autograd(stride, padding):
    torch.manual_seed(2)
    x = torch.randn(3, 4, 7, 6, dtype=torch.float64, requires_grad=True)
    s = (torch.rand(2, 4, dtype=torch.float64) * 5 - 2.5).requires_grad_()
    out = shift2d(x, s, stride, padding)
    og = torch.randn_like(out)
    out.backward(og)
    assert_close(oracle.shift2d_forward(x, s, stride, padding), out.detach().numpy(), 1e-12, "out")
    gin, gs = oracle.shift2d_backward(x, s, og, stride, padding, normalize_grad=False)
    assert_close(gin, x.grad.numpy(), 1e-12, "gin")
    assert_close(gs, s.grad.numpy(), 1e-11, "gshift")


def test_integer_rule_3d_known_answer():
    """Exact-integer T shift: gT = x[t+f+1] - x[t+f-1] (span 2, no 0.5), and the H / W gradients are
    evaluated on frame t+f-1 (cuda_src/rubiks3d_kernels.cu:290,359-431)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1, 5, 1, 4, 4))
    shift = np.array([[1.0], [0.25], [0.5]])
    og = np.zeros((1, 5, 1, 4, 4))
    og[0, 2, 0, 1, 1] = 1.0  # single upstream pixel: to=2, ho=1, wo=1
    gs = oracle.shift3d_backward_shift_raw(x, shift, og)
    rh, rw = 0.25, 0.5

    def bil(t, h0, w0):  # I2 with (rh, rw) on frame t
        p = lambda h, w: x[0, t, 0, h, w] if 0 <= t < 5 and 0 <= h < 4 and 0 <= w < 4 else 0.0
        return p(h0, w0) * (1 - rh) * (1 - rw) + p(h0, w0 + 1) * (1 - rh) * rw + \
            p(h0 + 1, w0) * rh * (1 - rw) + p(h0 + 1, w0 + 1) * rh * rw
    assert np.isclose(gs[0, 0], bil(4, 1, 1) - bil(2, 1, 1))  # frames to+f+1=4 and to+f-1=2
    p = lambda t, h, w: x[0, t, 0, h, w]
    gH = (1 - rw) * (p(2, 2, 1) - p(2, 1, 1)) + rw * (p(2, 2, 2) - p(2, 1, 2))  # on frame 2 = to+f-1
    assert np.isclose(gs[1, 0], gH)


def test_integer_rule_2d_known_answer():
    """2D integer H shift: 0.5 * central difference; W gradient stays on the floor row
    (cuda_src/rubiks2d_kernels.cu:189-253)."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 1, 5, 5))
    shift = np.array([[1.0], [0.5]])
    og = np.zeros((1, 1, 5, 5))
    og[0, 0, 2, 1] = 1.0
    _, gs = oracle.shift2d_backward(x, shift, og, 1, 0, normalize_grad=False)
    p = lambda h, w: x[0, 0, h, w] if 0 <= h < 5 and 0 <= w < 5 else 0.0
    h0, w0, rw = 3, 1, 0.5
    assert np.isclose(gs[0, 0], 0.5 * ((1 - rw) * (p(h0 + 1, w0) - p(h0 - 1, w0)) + rw * (p(h0 + 1, w0 + 1) - p(h0 - 1, w0 + 1))))
    assert np.isclose(gs[1, 0], p(h0, w0 + 1) - p(h0, w0))


@pytest.mark.parametrize("kind", ["rand3", "integer", "halves"])
@pytest.mark.parametrize("stride", [(1, 1, 1), (1, 2, 2)])
def test_oracle_adjoint_identity(kind, stride):
    """<shift(x), og> == <x, input_grad(og)> for every shift value (the input gradient is the exact adjoint)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4, 6, 8, 9))
    s = make_shift(rng, kind, 3, 6, np.float64)
    out = oracle.shift3d_forward(x, s, stride, 0)
    og = rng.standard_normal(out.shape)
    gin, _ = oracle.shift3d_backward(x, s, og, stride, 0, normalize_grad=False)
    assert np.isclose((out * og).sum(), (x * gin).sum(), rtol=1e-12)


def test_quantize_rounding_rules():
    """3D forward rounds the remainder half-up; backward rounds the NEGATED shift half-up, so the two
    disagree at exactly .5 (SURVEY.md section 7 hard parts); 2D rounds half away from zero."""
    x = np.arange(7, dtype=np.float64).reshape(1, 1, 1, 1, 7) + 1
    s = np.array([[0.0], [0.0], [0.5]])
    out = oracle.shift3d_forward(x, s, 1, 0, quantize=True)
    np.testing.assert_array_equal(out[0, 0, 0, 0], [2, 3, 4, 5, 6, 7, 0])   # reads w+1
    gin, _ = oracle.shift3d_backward(x, s, x, 1, 0, normalize_grad=False, quantize=True)
    np.testing.assert_array_equal(gin[0, 0, 0, 0], [1, 2, 3, 4, 5, 6, 7])   # -0.5 -> floor -1, r .5 -> 0
    x2 = x.reshape(1, 1, 1, 7)
    out2 = oracle.shift2d_forward(x2, np.array([[0.0], [-0.5]]), 1, 0, quantize=True)
    np.testing.assert_array_equal(out2[0, 0, 0], [0, 2, 3, 4, 5, 6, 7])     # round_fast(-0.5) = -1 (OOB), (0.5) = 1, (1.5) = 2


def test_normalize_rules():
    g = np.array([[3.0, 0.0, -2.0], [4.0, 0.0, 1.0], [0.0, 0.0, 2.0]])
    n = oracle.normalize_shift_grad_3d(g, 1.0)
    assert np.allclose(n[:, 0], [0.6, 0.8, 0.0]) and np.allclose(n[:, 1], 0) and np.allclose(np.linalg.norm(n[:, 2]), 1)
    n = oracle.normalize_shift_grad_3d(g, -1.0)
    assert np.allclose(n[:, 0], [1, 0, 0]) and np.allclose(n[:, 2], [-1, 0, 0])
    assert np.allclose(n[:, 1], g[:, 1])  # zero norm: nothing rewritten
    n = oracle.normalize_shift_grad_3d(g, 0.5)
    v = np.array([1.5, 4.0, 0.0])
    assert np.allclose(n[:, 0], v / np.linalg.norm(v))


def test_attention_oracle_matches_reference_module(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention_golden.npz"))
    for name in ("small", "sharp", "t1"):
        n, t, temp = g[name + "/meta"]
        out = oracle.attention_shift_forward(g[name + "/x"], g[name + "/weight"], temp, int(t))
        gx, gw = oracle.attention_shift_backward(g[name + "/x"], g[name + "/weight"], temp, int(t), g[name + "/og"])
        assert_close(out, g[name + "/out"], 1e-12, name)
        assert_close(gx, g[name + "/gx"], 1e-12, name)
        assert_close(gw, g[name + "/gweight"], 1e-9, name)


def test_empty_and_degenerate_shapes():
    x = np.zeros((0, 4, 3, 5, 5), dtype=np.float32)
    s = np.zeros((3, 3), dtype=np.float32)
    assert oracle.shift3d_forward(x, s).shape == (0, 4, 3, 5, 5)
    x = np.ones((1, 1, 1, 1, 1), dtype=np.float32)
    out = oracle.shift3d_forward(x, np.array([[0.5], [0.0], [0.0]], dtype=np.float32))
    assert out.item() == 0.5
