"""BASELINE C1: RubiksShift3D single layer, N=2 C=64 T=8 H=W=56, fp32, stride 1, no padding -- forward and backward of
the CUDA kernels (through the C ABI) beside the CPU oracle (oracle/, the C restatement of the reference's algorithm; the
reference itself has no CPU path, rubiksnet/shiftlib/rubiks3d/primitive.py:61) on 1 thread and on all host cores.

    python tools/bench_c1.py [--iters 100]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import oracle  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_backward, rubiks_shift_3d_forward  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=100)
    a = ap.parse_args()
    torch.manual_seed(0)
    x = torch.randn(2, 8, 64, 56, 56)
    shift = torch.rand(3, 64) * 2 - 1
    og = torch.randn(2, 8, 64, 56, 56)
    elems = x.numel()
    fwd_bytes, bwd_bytes = 2 * elems * 4, 3 * elems * 4
    print("C1: x [2,8,64,56,56] fp32 (%d elements), algorithmic bytes fwd %d / bwd %d" % (elems, fwd_bytes, bwd_bytes))
    xn, sn, gn = x.numpy(), shift.numpy(), og.numpy()
    cores = os.cpu_count()
    for threads in (1, cores):
        oracle.set_num_threads(threads)
        oracle.shift3d_forward(xn, sn, 1, 0)
        reps = 3 if threads == 1 else 10
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.shift3d_forward(xn, sn, 1, 0)
        tf = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.shift3d_backward(xn, sn, gn, 1, 0, normalize_grad=True)
        tb = (time.perf_counter() - t0) / reps
        print("CPU oracle (C port), %2d thread(s): fwd %8.3f ms %6.2f GB/s | bwd %8.3f ms %6.2f GB/s"
              % (threads, tf * 1e3, fwd_bytes / tf / 1e9, tb * 1e3, bwd_bytes / tb / 1e9))
    if torch.cuda.is_available():
        xd, sd, gd = x.cuda(), shift.cuda(), og.cuda()
        out = rubiks_shift_3d_forward(xd, sd, (1, 1, 1), 0)
        gin, gs = torch.empty_like(xd), torch.empty_like(sd)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        res = {}
        for name, fn in (("fwd", lambda: rubiks_shift_3d_forward(xd, sd, (1, 1, 1), 0, output=out)),
                         ("bwd", lambda: rubiks_shift_3d_backward(gd, xd, sd, (1, 1, 1), 0, True, x_grad_output=gin,
                                                                  shift_grad_output=gs))):
            for _ in range(5):
                fn()
            ts = []
            for _ in range(a.iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[name] = float(np.median(ts))
        print("librubiks_b200 on %s (L2 flushed, median of %d): fwd %.4f ms %.0f GB/s | bwd %.4f ms %.0f GB/s"
              % (torch.cuda.get_device_name(0), a.iters, res["fwd"], fwd_bytes / res["fwd"] / 1e6, res["bwd"],
                 bwd_bytes / res["bwd"] / 1e6))
        want = oracle.shift3d_forward(xn, sn, 1, 0)
        err = np.abs(out.cpu().numpy() - want).max()
        print("max |cuda - oracle| forward: %.3g (tolerance 1e-4)" % err)


if __name__ == "__main__":
    main()
