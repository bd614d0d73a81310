"""Shared test helpers (test infrastructure)."""
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")

# tolerances stated by BASELINE.json north_star: 1e-4 fp32, 1e-2 bf16 (also used for fp16); float64 1e-10
TOL = {"float32": 1e-4, "float64": 1e-10, "float16": 1e-2, "bfloat16": 1e-2}


def make_shift(rng, kind, dims, C, dtype=np.float32):
    if kind == "rand1":
        s = rng.uniform(-1, 1, size=(dims, C))
    elif kind == "rand3":
        s = rng.uniform(-3, 3, size=(dims, C))
    elif kind == "rand15":
        s = rng.uniform(-15, 15, size=(dims, C))
    elif kind == "integer":
        vals = np.array([-2.0, -1.0, 0.0, 1.0, 2.0, 0.25, -0.75, 1.5])
        s = vals[rng.integers(0, len(vals), size=(dims, C))]
        s[:, 0] = 0.0
        if C > 1:
            s[:, 1] = np.array([1.0, -1.0, 2.0])[:dims]
    elif kind == "halves":
        vals = np.array([-1.5, -0.5, 0.5, 1.5, 0.25, -0.25, 1.0, 0.0])
        s = vals[rng.integers(0, len(vals), size=(dims, C))]
    elif kind == "zero":
        s = np.zeros((dims, C))
    else:
        raise ValueError(kind)
    return s.astype(dtype)


def assert_close(actual, expected, tol, what=""):
    """|actual - expected| <= tol * max(1, |expected|) ELEMENT BY ELEMENT (absolute `tol` for values below 1, relative above:
    the meaning of north_star's "within 1e-4 fp32 / 1e-2 bf16").

    Reductions over many terms -- the shift gradients, recognised by "shift_grad" / "gshift" / "gweight" / "taps grad" in
    `what` -- are compared with `tol` relative to the LARGEST expected entry instead: an fp32 sum of 10^4..10^6 products has a
    rounding error proportional to the sum of the |terms|, not to the (possibly cancelling) result of one channel, and the
    reference itself adds them with atomics in arbitrary order."""
    actual = np.asarray(actual, dtype=np.float64)
    expected = np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    if actual.size == 0:
        return
    err = np.abs(actual - expected)
    reduction = any(k in what for k in ("shift_grad", "gshift", "gweight", "taps grad", "gs"))
    if reduction:
        scale = max(1.0, float(np.abs(expected).max()))
        assert float(err.max()) <= tol * scale, "%s: max abs err %.3e > %.1e * %.3g" % (what, float(err.max()), tol, scale)
        return
    bound = tol * np.maximum(1.0, np.abs(expected))
    bad = err > bound
    if bad.any():
        i = int(np.argmax(err - bound))
        raise AssertionError("%s: %d of %d elements off; worst |%.6g - %.6g| = %.3e > %.1e * max(1, |expected|)" % (
            what, int(bad.sum()), err.size, actual.flat[i], expected.flat[i], float(err.flat[i]), tol))


def load_golden(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        return None
    return np.load(path)
