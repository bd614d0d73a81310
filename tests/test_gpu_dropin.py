"""Boundary proof: the reference's OWN `rubiksnet` Python package, unmodified, runs on
librubiks_b200 when `rubiksnet_cuda` resolves to this repo's stand-in (dropin/rubiksnet_cuda.py ->
rubiksnet_b200/rubiksnet_cuda.py -> C ABI), and gives the numbers of the reference's CUDA extension.

Recipe = the reference's scripts/test_installation.py:6-10 (RubiksNet-Large, random init, seed 0) plus one backward.
Each arm runs in its own subprocess because both native modules want the name `rubiksnet_cuda`."""
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import REPO  # noqa: E402

pytestmark = pytest.mark.gpu

_REF = os.path.join(REPO, "baseline", "_ref")


def _run(native, out, tier="large"):
    env = dict(os.environ)
    env.pop("PYTHONPATH", None)
    r = subprocess.run([sys.executable, os.path.join(REPO, "tests", "dropin_check.py"), "--native", native, "--tier", tier,
                        "--out", out], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, "dropin_check --native %s failed:\n%s\n%s" % (native, r.stdout[-2000:], r.stderr[-4000:])
    return np.load(out)


@pytest.mark.skipif(not os.path.isdir(os.path.join(_REF, "rubiksnet")), reason="baseline/_ref not on this box")
def test_reference_package_runs_on_librubiks_b200(tmp_path):
    ours = _run("b200", str(tmp_path / "b200.npz"))
    assert str(ours["impl"]) == "librubiks_b200"
    assert ours["prediction"].shape == (2, 42) and np.isfinite(ours["prediction"]).all()
    ref = _run("reference", str(tmp_path / "ref.npz"))
    assert str(ref["impl"]) != "librubiks_b200"
    # same Python, same cuDNN convolutions, same seed: the only difference is the native shift module (fp32).
    scale = max(1.0, float(np.abs(ref["prediction"]).max()))
    assert float(np.abs(ours["prediction"] - ref["prediction"]).max()) <= 1e-4 * scale
    assert abs(float(ours["loss"]) - float(ref["loss"])) <= 1e-4 * max(1.0, abs(float(ref["loss"])))
    for k in ref.files:
        if not k.startswith("grad:"):
            continue
        a, b = ours[k].astype(np.float64), ref[k].astype(np.float64)
        if k.endswith("shift"):
            # unit-normalised per-channel gradients; the reference sums with fp32 atomics in arbitrary order and channels
            # with a tiny raw gradient amplify that noise -> mean over channels tight, max loose (as in test_gpu_block)
            d = np.abs(a - b)  # measured on RubiksNet-Large (51 shift layers): mean 2.4e-3, max 1.4e-2
            assert d.mean() <= 5e-3 and d.max() <= 0.2, (k, d.mean(), d.max())
        else:
            # fp32 end to end; the two native modules sum the input gradient of the shift in a different order and 51
            # training-mode BatchNorm layers amplify that on the way down to conv1 (measured 4.7e-3 there): 1e-2 of max|grad|
            assert np.abs(a - b).max() <= 1e-2 * max(1e-6, np.abs(b).max()), k
