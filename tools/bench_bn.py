"""Op-level benchmark of the BatchNorm(+ReLU) passes on every RubiksNet-Large block geometry (bf16, 32 clips):
forward statistics, statistics + apply, backward (reduce + apply), with and without an L2 flush between calls (the
14x14 / 7x7 tensors of a step are L2-resident when the next kernel reads them).

    python tools/bench_bn.py [--batch 32] [--iters 20] [--only layer3.x]
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import ops  # noqa: E402

LAYERS = [("layer0", 72, 112), ("layer1.x", 72, 56), ("layer2.x", 144, 28), ("layer3.x", 288, 14), ("layer4.x", 576, 7)]
_flush = None


def timeit(fn, iters, flush):
    global _flush
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            if _flush is None:
                _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    print("device:", torch.cuda.get_device_name(0))
    for name, c, h in LAYERS:
        if a.only and a.only != name:
            continue
        ni = a.batch * 8
        x = torch.randn(ni, c, h, h, device="cuda").bfloat16()
        dy = torch.randn_like(x)
        res = torch.randn_like(x)
        g, b = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        _, mi, sb = ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=True, apply=False)
        unit = x.numel() * 2
        cases = {
            "stats": (lambda: ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=True, apply=False), unit),
            "stats+apply": (lambda: ops.bn_forward(x, g, b, rm, rv, True, 0.1, 1e-5, relu=True, apply=True), 3 * unit),
            "apply": (lambda: ops.bn_apply(x, sb, relu=True), 2 * unit),
            "bwd": (lambda: ops.bn_backward(x, dy, None, g, mi, sb, True, relu=True), 5 * unit),
            "bwd+res": (lambda: ops.bn_backward(x, dy, res, g, mi, sb, True, relu=True), 6 * unit),
        }
        line = "%-9s C=%-3d H=%-3d unit %.1f MB |" % (name, c, h, unit / 1e6)
        for cname, (fn, nbytes) in cases.items():
            for flush in (True, False):
                ms = timeit(fn, a.iters, flush)
                line += " %s%s %.3f ms %.0f GB/s |" % (cname, "" if flush else "(L2 warm)", ms, nbytes / ms / 1e6)
        print(line, flush=True)


if __name__ == "__main__":
    main()
