"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs; against the reference's recorded outputs (tests/golden); against the live reference
extension when baseline/_ref travelled to the box; and size-independent properties at BASELINE sizes.
Tolerances (BASELINE.json north_star): 1e-4 fp32, 1e-2 bf16/fp16; 1e-10 float64; bit-exact for the
quantize (pure copy) paths."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
from golden.make_golden import cases2d, cases3d  # noqa: E402
from helpers import REPO, TOL, assert_close, load_golden, make_shift  # noqa: E402

from rubiksnet_b200 import _lib  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks2d.primitive import rubiks2d, rubiks2d_backward, rubiks2d_forward  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks3d.primitive import (rubiks_shift_3d, rubiks_shift_3d_backward,  # noqa: E402
                                                        rubiks_shift_3d_forward)

pytestmark = pytest.mark.gpu
GOLD = load_golden("shift_golden.npz")
TDT = {"float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "bfloat16": torch.bfloat16}


@pytest.fixture(autouse=True)
def _auto_impl():
    _lib.set_impl(_lib.RB_IMPL_AUTO)
    yield
    _lib.set_impl(_lib.RB_IMPL_AUTO)


def _cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def _run3d(x, shift, og, stride, padding, dtype, normalize_grad=False, t_factor=1.0, quantize=False):
    """CUDA path on (x, shift, og) given as float32/float64 numpy; returns numpy float64 results.
    For 16-bit dtypes the inputs are rounded to that dtype first and the oracle sees the rounded values."""
    td = TDT[dtype]
    tx, tg = _cuda(x, td), _cuda(og, td)
    ts = _cuda(shift) if dtype in ("float16", "bfloat16") else _cuda(shift, td)
    out = rubiks_shift_3d_forward(tx, ts, stride, padding, quantize=quantize)
    impl_f = _lib.last_impl()
    gin, gs = rubiks_shift_3d_backward(tg, tx, ts, stride, padding, normalize_grad, t_factor, quantize)
    impl_b = _lib.last_impl()
    torch.cuda.synchronize()
    rx, rg = tx.double().cpu().numpy(), tg.double().cpu().numpy()
    return (out.double().cpu().numpy(), gin.double().cpu().numpy(), gs.double().cpu().numpy(), rx, rg, impl_f, impl_b)


def _oracle3d(rx, shift, rg, stride, padding, dtype, normalize_grad, t_factor, quantize):
    odt = np.float64 if dtype == "float64" else np.float32
    o = oracle.shift3d_forward(rx.astype(odt), shift.astype(odt), stride, padding, quantize)
    gin, gs = oracle.shift3d_backward(rx.astype(odt), shift.astype(odt), rg.astype(odt), stride, padding,
                                      normalize_grad, t_factor, quantize)
    return o, gin, gs


SHAPES = [
    # (N,T,C,H,W), stride, padding
    ((2, 8, 12, 14, 14), (1, 1, 1), (0, 0, 0)),     # small-map regime: several channels per CTA
    ((2, 8, 5, 28, 28), (1, 1, 1), (0, 0, 0)),
    ((1, 8, 3, 56, 56), (1, 1, 1), (0, 0, 0)),      # row-band regime
    ((1, 4, 2, 112, 112), (1, 1, 1), (0, 0, 0)),
    ((2, 8, 6, 28, 28), (1, 2, 2), (0, 0, 0)),      # down-sampling blocks
    ((1, 8, 3, 56, 56), (1, 2, 2), (0, 0, 0)),
    ((2, 8, 9, 7, 7), (1, 1, 1), (0, 0, 0)),
    ((1, 3, 4, 9, 11), (1, 1, 1), (0, 0, 0)),       # ragged / odd extents
    ((2, 5, 3, 13, 10), (1, 2, 2), (0, 0, 0)),
    ((2, 5, 4, 8, 9), (2, 2, 1), (1, 1, 0)),        # generic-kernel geometry (T stride, padding)
    ((1, 20, 2, 6, 6), (1, 1, 1), (0, 0, 0)),       # more frames than the tiled kernels keep resident
]


@pytest.mark.parametrize("impl", ["auto", "tiled", "generic"])
@pytest.mark.parametrize("kind", ["rand1", "rand3", "rand15", "integer", "halves", "zero"])
@pytest.mark.parametrize("shape,stride,padding", SHAPES)
def test_shift3d_fp32_vs_oracle(shape, stride, padding, kind, impl):
    rng = np.random.default_rng(sum(shape) * 31 + sum(stride) * 7 + len(kind))
    fast_geometry = stride[0] == 1 and stride[1] == stride[2] and stride[1] in (1, 2) and padding == (0, 0, 0) and shape[1] <= 16
    if impl == "generic":
        _lib.set_impl(_lib.RB_IMPL_GENERIC)
    elif impl == "tiled":
        if not fast_geometry:
            pytest.skip("geometry not covered by the tiled kernels")
        _lib.set_impl(_lib.RB_IMPL_TILED)
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, kind, 3, shape[2])
    o_ref = oracle.shift3d_forward(x, s, stride, padding)
    og = rng.standard_normal(o_ref.shape).astype(np.float32)
    out, gin, gs, rx, rg, impl_f, impl_b = _run3d(x, s, og, stride, padding, "float32")
    _, gin_ref, gs_ref = _oracle3d(rx, s, rg, stride, padding, "float32", False, 1.0, False)
    assert_close(out, o_ref, TOL["float32"], "out")
    assert_close(gin, gin_ref, TOL["float32"], "x_grad")
    assert_close(gs, gs_ref, TOL["float32"], "shift_grad")
    if impl == "auto" and fast_geometry:
        want = _lib.RB_IMPL_STRIP if stride[1] == 1 else _lib.RB_IMPL_TILED
        assert impl_f == want and impl_b == want, "TMA-staged sm_100a path was not taken"


@pytest.mark.parametrize("dtype", ["bfloat16", "float16", "float64"])
@pytest.mark.parametrize("kind", ["rand1", "integer"])
@pytest.mark.parametrize("shape,stride,padding", SHAPES[:6] + SHAPES[9:10])
def test_shift3d_other_dtypes_vs_oracle(shape, stride, padding, kind, dtype):
    rng = np.random.default_rng(7)
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, kind, 3, shape[2])
    oshape = oracle.shift3d_forward(x, s, stride, padding).shape
    og = rng.standard_normal(oshape).astype(np.float32)
    out, gin, gs, rx, rg, _, _ = _run3d(x, s, og, stride, padding, dtype)
    o_ref, gin_ref, gs_ref = _oracle3d(rx, s, rg, stride, padding, dtype, False, 1.0, False)
    assert_close(out, o_ref, TOL[dtype], "out")
    assert_close(gin, gin_ref, TOL[dtype], "x_grad")
    # the shift gradient accumulates in fp32 from exact 16-bit inputs: only the output rounding differs
    assert_close(gs, gs_ref, 1e-4 if dtype != "float64" else 1e-10, "shift_grad")


@pytest.mark.parametrize("t_factor", [1.0, 0.5, -1.0, 8 / 14])
def test_shift3d_normalised_grad(t_factor):
    rng = np.random.default_rng(11)
    shape, stride, padding = (2, 8, 12, 14, 14), (1, 1, 1), (0, 0, 0)
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, "rand1", 3, 12)
    og = rng.standard_normal(shape).astype(np.float32)
    _, _, gs, rx, rg, _, _ = _run3d(x, s, og, stride, padding, "float32", True, t_factor)
    _, _, gs_ref = _oracle3d(rx, s, rg, stride, padding, "float32", True, t_factor, False)
    assert_close(gs, gs_ref, 1e-4, "normalised shift_grad")
    if t_factor > 0:
        v = gs.copy()
        np.testing.assert_allclose(np.linalg.norm(v, axis=0), 1.0, rtol=1e-5)


@pytest.mark.parametrize("stride", [(1, 1, 1), (1, 2, 2), (2, 1, 2)])
@pytest.mark.parametrize("kind", ["halves", "rand3"])
def test_shift3d_quantize_bit_exact(stride, kind):
    rng = np.random.default_rng(5)
    shape = (2, 6, 8, 10, 12)
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, kind, 3, 8)
    o_ref = oracle.shift3d_forward(x, s, stride, 0, quantize=True)
    og = rng.standard_normal(o_ref.shape).astype(np.float32)
    out, gin, gs, rx, rg, _, _ = _run3d(x, s, og, stride, (0, 0, 0), "float32", quantize=True)
    gin_ref, gs_ref = oracle.shift3d_backward(x, s, og, stride, 0, False, 1.0, True)
    np.testing.assert_array_equal(out.astype(np.float32), o_ref)
    np.testing.assert_array_equal(gin.astype(np.float32), gin_ref)
    assert_close(gs, gs_ref, 1e-4, "shift_grad (never quantised)")


@pytest.mark.skipif(GOLD is None, reason="tests/golden/shift_golden.npz not generated yet")
@pytest.mark.parametrize("name", sorted(cases3d()))
def test_shift3d_vs_reference_golden(name):
    cfg = cases3d()[name]
    g = lambda k: GOLD["3d/%s/%s" % (name, k)]
    dtype = cfg.get("dtype", "float32")
    out, gin, gs, _, _, _, _ = _run3d(g("x"), g("shift"), g("og"), cfg["stride"], cfg["padding"], dtype,
                                      cfg.get("normalize_grad", True), cfg.get("t_factor", 1.0),
                                      cfg.get("quantize", False))
    assert_close(out, g("out"), TOL[dtype], "out")
    assert_close(gin, g("gin"), TOL[dtype], "x_grad")
    assert_close(gs, g("gshift"), 1e-4 if dtype == "float32" else 1e-9, "shift_grad")


# ------------------------------------------------------------------------------------------ 2D

SHAPES2D = [((4, 12, 14, 14), (1, 1), (0, 0)), ((3, 5, 28, 28), (2, 2), (0, 0)), ((2, 3, 57, 56), (1, 1), (0, 0)),
            ((2, 4, 8, 9), (2, 1), (1, 0)), ((2, 4, 9, 11), (3, 2), (2, 1))]


@pytest.mark.parametrize("dtype", ["float32", "bfloat16", "float16", "float64"])
@pytest.mark.parametrize("kind", ["rand1", "rand3", "integer", "zero"])
@pytest.mark.parametrize("shape,stride,padding", SHAPES2D)
def test_shift2d_vs_oracle(shape, stride, padding, kind, dtype):
    rng = np.random.default_rng(13)
    td = TDT[dtype]
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, kind, 2, shape[1])
    oshape = oracle.shift2d_forward(x, s, stride, padding).shape
    og = rng.standard_normal(oshape).astype(np.float32)
    tx, tg = _cuda(x, td), _cuda(og, td)
    ts = _cuda(s) if dtype in ("float16", "bfloat16") else _cuda(s, td)
    out = rubiks2d_forward(tx, ts, stride, padding)
    gin, gs = rubiks2d_backward(tg, tx, ts, stride, padding, normalize_grad=False)
    odt = np.float64 if dtype == "float64" else np.float32
    rx, rg = tx.double().cpu().numpy().astype(odt), tg.double().cpu().numpy().astype(odt)
    o_ref = oracle.shift2d_forward(rx, s.astype(odt), stride, padding)
    gin_ref, gs_ref = oracle.shift2d_backward(rx, s.astype(odt), rg, stride, padding, normalize_grad=False)
    assert_close(out.double().cpu().numpy(), o_ref, TOL[dtype], "out")
    assert_close(gin.double().cpu().numpy(), gin_ref, TOL[dtype], "x_grad")
    assert_close(gs.double().cpu().numpy(), gs_ref, 1e-4 if dtype != "float64" else 1e-10, "shift_grad")


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
@pytest.mark.parametrize("kind", ["rand1", "rand3", "integer", "halves", "zero"])
@pytest.mark.parametrize("shape,stride", [((16, 12, 28, 28), 2), ((8, 5, 56, 56), 2), ((6, 7, 14, 14), 1), ((3, 4, 30, 22), 2)])
def test_shift2d_tiled_kernel_vs_oracle(shape, stride, kind, dtype):
    """2D shift on the TMA-staged tiled kernel (stride 2: the down-sampling blocks of the attention-quantized variant;
    groups of images share a CTA, 2D integer rule in the slow path) vs the oracle; the dispatcher must pick it."""
    rng = np.random.default_rng(17)
    td = TDT[dtype]
    x = rng.standard_normal(shape).astype(np.float32)
    s = make_shift(rng, kind, 2, shape[1])
    og = rng.standard_normal(oracle.shift2d_forward(x, s, stride, 0).shape).astype(np.float32)
    tx, tg, ts = _cuda(x, td), _cuda(og, td), _cuda(s)
    try:
        _lib.set_impl(_lib.RB_IMPL_TILED if stride == 1 else _lib.RB_IMPL_AUTO)
        out = rubiks2d_forward(tx, ts, stride, 0)
        assert _lib.last_impl() == _lib.RB_IMPL_TILED
        gin, gs = rubiks2d_backward(tg, tx, ts, stride, 0, normalize_grad=False)
        assert _lib.last_impl() == _lib.RB_IMPL_TILED
    finally:
        _lib.set_impl(_lib.RB_IMPL_AUTO)
    rx, rg = tx.float().cpu().numpy(), tg.float().cpu().numpy()
    o_ref = oracle.shift2d_forward(rx, s, stride, 0)
    gin_ref, gs_ref = oracle.shift2d_backward(rx, s, rg, stride, 0, normalize_grad=False)
    assert_close(out.float().cpu().numpy(), o_ref, TOL[dtype], "out")
    assert_close(gin.float().cpu().numpy(), gin_ref, TOL[dtype], "x_grad")
    assert_close(gs.float().cpu().numpy(), gs_ref, 1e-4, "shift_grad")


@pytest.mark.skipif(GOLD is None, reason="tests/golden/shift_golden.npz not generated yet")
@pytest.mark.parametrize("name", sorted(cases2d()))
def test_shift2d_vs_reference_golden(name):
    cfg = cases2d()[name]
    g = lambda k: GOLD["2d/%s/%s" % (name, k)]
    tx, ts, tg = _cuda(g("x")), _cuda(g("shift")), _cuda(g("og"))
    q = cfg.get("quantize", False)
    out = rubiks2d_forward(tx, ts, cfg["stride"], cfg["padding"], quantize=q)
    gin, gs = rubiks2d_backward(tg, tx, ts, cfg["stride"], cfg["padding"], cfg.get("normalize_grad", True),
                                cfg.get("enable_shift_grad", True), q)
    tol = 1e-10 if cfg.get("dtype") == "float64" else 1e-4
    assert_close(out.cpu().numpy(), g("out"), tol, "out")
    assert_close(gin.cpu().numpy(), g("gin"), tol, "x_grad")
    assert_close(gs.cpu().numpy(), g("gshift"), tol, "shift_grad")


# --------------------------------------------------------------------- properties at full size


def test_full_size_properties_c1():
    """BASELINE config C1 (N=2,C=64,T=8,H=W=56): adjoint identity, zero-shift identity (bit exact) and
    integer-shift = pure translation (bit exact), linearity in x."""
    torch.manual_seed(0)
    x = torch.randn(2, 8, 64, 56, 56, device="cuda")
    shift = torch.rand(3, 64, device="cuda") * 2 - 1
    out = rubiks_shift_3d_forward(x, shift, 1, 0)
    og = torch.randn_like(out)
    gin, _ = rubiks_shift_3d_backward(og, x, shift, 1, 0, False)
    lhs, rhs = (out.double() * og.double()).sum().item(), (x.double() * gin.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0)
    zero = torch.zeros(3, 64, device="cuda")
    assert torch.equal(rubiks_shift_3d_forward(x, zero, 1, 0), x)
    ints = torch.zeros(3, 64, device="cuda")
    ints[0], ints[1], ints[2] = 1.0, -2.0, 3.0
    moved = rubiks_shift_3d_forward(x, ints, 1, 0)
    expect = torch.zeros_like(x)
    expect[:, :-1, :, 2:, :-3] = x[:, 1:, :, :-2, 3:]
    assert torch.equal(moved, expect)
    y = torch.randn_like(x)
    lin = rubiks_shift_3d_forward(2 * x + y, shift, 1, 0)
    assert torch.allclose(lin, 2 * out + rubiks_shift_3d_forward(y, shift, 1, 0), atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_large_layer_shapes_tiled_equals_generic(dtype):
    """Every shift-layer geometry of RubiksNet-Large at batch 2 clips: the tiled
    kernels and the generic gather kernels must agree."""
    torch.manual_seed(1)
    layers = [(72, 112, 1), (72, 112, 2), (72, 56, 1), (144, 56, 2), (144, 28, 1), (288, 28, 2), (288, 14, 1),
              (576, 14, 2), (576, 7, 1)]
    for C, H, S in layers:
        x = torch.randn(2, 8, C, H, H, device="cuda").to(dtype)
        shift = (torch.rand(3, C, device="cuda") * 2 - 1)
        shift[:, 0] = 0.0
        shift[0, 1] = 1.0
        res = {}
        for impl in (_lib.RB_IMPL_AUTO, _lib.RB_IMPL_TILED, _lib.RB_IMPL_GENERIC):
            _lib.set_impl(impl)
            out = rubiks_shift_3d_forward(x, shift, (1, S, S), 0)
            og = torch.randn(out.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(3)).to(dtype)
            gin, gs = rubiks_shift_3d_backward(og, x, shift, (1, S, S), 0, False)
            res[impl] = (out.float(), gin.float(), gs.float())
        tol = 1e-5 if dtype == torch.float32 else 1e-2
        for fast in (_lib.RB_IMPL_AUTO, _lib.RB_IMPL_TILED):
            for a, b, what in zip(res[fast], res[_lib.RB_IMPL_GENERIC], ("out", "x_grad", "shift_grad")):
                scale = max(1.0, b.abs().max().item())
                assert (a - b).abs().max().item() <= (1e-4 if what == "shift_grad" else tol) * scale, (C, H, S, what, fast)


def test_pretrained_shift_distributions(golden_dir):
    """The 51 trained shift tensors of ssv2_large (values up to |14|, a few exact integers)."""
    shifts = np.load(os.path.join(golden_dir, "pretrained_shifts_large.npz"))
    rng = np.random.default_rng(2)
    for key in list(shifts.keys())[::6]:
        s = shifts[key].astype(np.float32)
        C = s.shape[1]
        H = {72: 28, 144: 28, 288: 14, 576: 7}[C]
        x = rng.standard_normal((1, 8, C, H, H)).astype(np.float32)
        og = rng.standard_normal(x.shape).astype(np.float32)
        out, gin, gs, rx, rg, _, _ = _run3d(x, s, og, (1, 1, 1), (0, 0, 0), "float32")
        o_ref, gin_ref, gs_ref = _oracle3d(rx, s, rg, (1, 1, 1), (0, 0, 0), "float32", False, 1.0, False)
        assert_close(out, o_ref, 1e-4, key)
        assert_close(gin, gin_ref, 1e-4, key)
        assert_close(gs, gs_ref, 1e-4, key)


def test_autograd_function_and_empty():
    torch.manual_seed(0)
    x = torch.randn(2, 4, 6, 9, 9, device="cuda", requires_grad=True)
    shift = (torch.rand(3, 6, device="cuda") * 2 - 1).requires_grad_()
    out = rubiks_shift_3d(x, shift, stride=(1, 2, 2), normalize_grad=False)
    out.square().sum().backward()
    gin_ref, gs_ref = oracle.shift3d_backward(x.detach().cpu().numpy(), shift.detach().cpu().numpy(),
                                              (2 * out).detach().cpu().numpy(), (1, 2, 2), 0, normalize_grad=False)
    assert_close(x.grad.cpu().numpy(), gin_ref, 1e-4, "autograd x_grad")
    assert_close(shift.grad.cpu().numpy(), gs_ref, 1e-4, "autograd shift_grad")
    # "auto" temporal factor = T / H (primitive.py:209-210)
    shift.grad = None
    rubiks_shift_3d(x, shift, normalize_t_factor="auto").sum().backward()
    assert torch.isfinite(shift.grad).all()
    # empty batch
    e = torch.zeros(0, 4, 6, 9, 9, device="cuda")
    assert rubiks_shift_3d_forward(e, shift.detach(), 1, 0).shape == (0, 4, 6, 9, 9)
    x2 = torch.randn(3, 6, 9, 9, device="cuda", requires_grad=True)
    s2 = (torch.rand(2, 6, device="cuda") * 2 - 1).requires_grad_()
    rubiks2d(x2, s2, stride=2).sum().backward()
    assert x2.grad.shape == x2.shape and s2.grad.shape == s2.shape


def test_error_paths_on_gpu():
    x = torch.randn(2, 4, 6, 9, 9, device="cuda")
    shift = torch.zeros(3, 6, device="cuda")
    with pytest.raises(RuntimeError, match="contiguous"):
        from rubiksnet_b200 import rubiksnet_cuda
        rubiksnet_cuda.rubiks_shift_3d_forward_float(x.transpose(3, 4), shift, [1, 1, 1], [0, 0, 0], False, torch.empty_like(x))
    with pytest.raises(RuntimeError):
        rubiksnet_cuda.rubiks_shift_3d_forward_double(x, shift, [1, 1, 1], [0, 0, 0], False, torch.empty_like(x))
    with pytest.raises(RuntimeError, match="ShapeException"):
        x2 = x[:, 0].contiguous()
        rubiksnet_cuda.rubiks2d_forward(x2, torch.zeros(2, 5, device="cuda"), [1, 1], [0, 0], False, torch.empty_like(x2))
    with pytest.raises(_lib.RubiksCudaError, match="stride"):
        rubiksnet_cuda.rubiks_shift_3d_forward_float(x, shift, [0, 1, 1], [0, 0, 0], False, torch.empty_like(x))


def test_non_default_stream_and_launch_counter():
    x = torch.randn(2, 8, 16, 14, 14, device="cuda")
    shift = torch.rand(3, 16, device="cuda") - 0.5
    ref = rubiks_shift_3d_forward(x, shift, 1, 0)
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = rubiks_shift_3d_forward(x, shift, 1, 0)
    s.synchronize()
    assert torch.equal(out, ref)
    assert _lib.launch_count() == 1


# --------------------------------------------------------- live reference extension (if present)

def _reference_ext():
    ref = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "rubiksnet")):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import rubiksnet_cuda as ref_ext  # the reference's pybind module
        return ref_ext
    except Exception:
        return None


@pytest.mark.parametrize("stride", [(1, 1, 1), (1, 2, 2)])
@pytest.mark.parametrize("kind", ["rand1", "rand3", "integer"])
def test_live_reference_extension_3d(stride, kind):
    ref_ext = _reference_ext()
    if ref_ext is None:
        pytest.skip("baseline/_ref (reference CUDA extension) not available on this box")
    rng = np.random.default_rng(21)
    shape = (2, 8, 24, 28, 28)
    x = _cuda(rng.standard_normal(shape).astype(np.float32))
    s = _cuda(make_shift(rng, kind, 3, 24))
    out = rubiks_shift_3d_forward(x, s, stride, 0)
    og = torch.randn_like(out)
    gin, gs = rubiks_shift_3d_backward(og, x, s, stride, 0, True)
    r_out = torch.zeros_like(out)
    assert ref_ext.rubiks_shift_3d_forward_float(x, s, list(stride), [0, 0, 0], False, r_out) == 0
    r_gin, r_gs = torch.zeros_like(x), torch.zeros_like(s)
    assert ref_ext.rubiks_shift_3d_backward_float(x, s, og, list(stride), [0, 0, 0], r_gin, r_gs, True, 1.0, False) == 0
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), r_out.cpu().numpy(), 1e-4, "out vs reference ext")
    assert_close(gin.cpu().numpy(), r_gin.cpu().numpy(), 1e-4, "x_grad vs reference ext")
    assert_close(gs.cpu().numpy(), r_gs.detach().cpu().numpy(), 1e-4, "shift_grad vs reference ext")


def test_non_finite_inputs_strip_vs_reference_semantics():
    """the strip kernels give out-of-range COLUMN taps a zero weight instead of a predicate, so an
    inf / NaN that sits in the memory word such a tap is clamped onto (the last / first elements of the neighbouring row)
    turns 0 * inf into NaN where the reference semantics (generic kernel, oracle) give a finite value.
      * a non-finite value in the interior columns behaves exactly like the reference;
      * with a non-finite value in a border column, every position where the two kernels disagree about finiteness lies in
        a border column (within 2 of the edge) of the same channel -- bounded and documented, finite inputs are unaffected."""
    rng = np.random.default_rng(5)
    N, T, C, H, W = 1, 8, 4, 14, 14
    shift = torch.from_numpy(rng.uniform(-1, 1, size=(3, C)).astype(np.float32)).cuda()
    base = rng.standard_normal((N, T, C, H, W)).astype(np.float32)

    def run(x, impl):
        _lib.set_impl(impl)
        out = rubiks_shift_3d_forward(torch.from_numpy(x).cuda(), shift, (1, 1, 1), 0)
        assert _lib.last_impl() == (impl if impl != _lib.RB_IMPL_AUTO else _lib.RB_IMPL_STRIP)
        return out.cpu().numpy()

    x = base.copy()
    x[0, 3, 1, 6, 7] = np.inf                     # interior column
    a, b = run(x, _lib.RB_IMPL_AUTO), run(x, _lib.RB_IMPL_GENERIC)
    assert np.array_equal(np.isfinite(a), np.isfinite(b))
    m = np.isfinite(b)
    assert np.abs(a[m] - b[m]).max() <= 1e-4 * max(1.0, np.abs(b[m]).max())

    x = base.copy()
    x[0, 3, 2, 6, W - 1] = np.inf                 # last column of a row: neighbour of the next row's out-of-range taps
    a, b = run(x, _lib.RB_IMPL_AUTO), run(x, _lib.RB_IMPL_GENERIC)
    diff = np.argwhere(np.isfinite(a) != np.isfinite(b))
    for n, t, c, h, w in diff:
        assert c == 2 and (w <= 1 or w >= W - 2), (n, t, c, h, w)
    m = np.isfinite(a) & np.isfinite(b)
    assert np.abs(a[m] - b[m]).max() <= 1e-4 * max(1.0, np.abs(b[m]).max())
