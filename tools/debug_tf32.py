"""Error structure of the tf32 pointwise conv against fp64 (debug aid for csrc/pw_conv_tf32.cu)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rubiksnet_b200 import ops

torch.manual_seed(0)
for (ni, k, n, h, w_) in [(1, 16, 8, 8, 16), (2, 54, 54, 16, 16), (4, 54, 54, 112, 112)]:
    x = torch.randn(ni, k, h, w_, device="cuda")
    w = torch.randn(n, k, device="cuda") / k ** 0.5
    got = ops.pw_conv_f32(x, w)
    want = torch.einsum("nk,ikhw->inhw", w.double(), x.double())
    err = (got.double() - want).abs()
    print("geom", (ni, k, n, h, w_), "max err %.3g of %.3g; zeros in got: %.3f; nan: %d" % (
        err.max().item(), want.abs().max().item(), (got == 0).float().mean().item(), int(torch.isnan(got).sum())))
    if err.max().item() > 1e-2:
        bad = err > 1e-2
        print("  bad fraction %.3f; by channel:" % bad.float().mean().item(), bad.float().mean(dim=(0, 2, 3))[:16].tolist())
        flat = bad[0].reshape(n, -1).float()
        print("  by position block of 32 (image 0):", flat.mean(0).reshape(-1, 32).mean(1)[:16].tolist())
        # which single input channel / permutation explains it? probe with one-hot weights
        w1 = torch.zeros(n, k, device="cuda"); w1[0, 0] = 1.0
        g1 = ops.pw_conv_f32(x, w1)
        print("  one-hot W[0,0]: out[0,0,:8]", g1[0, 0].flatten()[:8].tolist(), "x[0,0,:8]", x[0, 0].flatten()[:8].tolist())
        for kk in range(min(k, 8)):
            w1 = torch.zeros(n, k, device="cuda"); w1[0, kk] = 1.0
            g1 = ops.pw_conv_f32(x, w1)[0, 0].flatten()
            # find which (channel, position) of x row matches g1[0]
            match = (x[0].reshape(k, -1) - g1[None, :]).abs().mean(1)
            print("   W[0,%d] -> output row 0 matches x channel %d (mean abs diff %.3g)" % (kk, int(match.argmin()), match.min().item()))
