"""Micro-benchmark of the 1x1-convolution formulations on the RubiksNet-Large layer shapes (bf16, NCHW):
 (a) fused.conv1x1 = batched GEMM (cuBLAS)  (b) F.conv2d (cuDNN)  -- forward and forward+backward."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rubiksnet_b200 import fused  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    torch.backends.cudnn.benchmark = True
    NI = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    for C, H in [(72, 112), (72, 56), (144, 28), (288, 14), (576, 7)]:
        x = torch.randn(NI, C, H, H, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        w = (torch.randn(C, C, 1, 1, device="cuda") * 0.05).requires_grad_()
        g = torch.randn(NI, C, H, H, device="cuda", dtype=torch.bfloat16)
        wb = w.detach().bfloat16()
        res = {}
        res["bmm fwd"] = timeit(lambda: fused.conv1x1(x.detach(), w.detach()))
        res["cudnn fwd"] = timeit(lambda: F.conv2d(x.detach(), wb))

        def fb_bmm():
            out = fused.conv1x1(x, w)
            torch.autograd.grad(out, (x, w), g)

        def fb_cudnn():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = F.conv2d(x, w)
            torch.autograd.grad(out, (x, w), g)

        res["bmm f+b"] = timeit(fb_bmm)
        res["cudnn f+b"] = timeit(fb_cudnn)
        xl = x.detach().contiguous(memory_format=torch.channels_last).requires_grad_()
        gl = g.contiguous(memory_format=torch.channels_last)

        def fb_cl():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = F.conv2d(xl, w)
            torch.autograd.grad(out, (xl, w), gl)

        res["cudnn-NHWC f+b"] = timeit(fb_cl)
        mb = x.numel() * 2 / 1e6
        ideal = 9 * mb / 6.4e3  # fwd 2 passes, dgrad 2, wgrad 2 (+3 for re-reads): ~9 tensor passes at 6.4 TB/s, in ms
        print("C=%3d H=%3d tensor %.1f MB | " % (C, H, mb) + " | ".join("%s %.3f ms" % kv for kv in res.items()) +
              " | ~roofline f+b %.3f ms" % ideal, flush=True)


if __name__ == "__main__":
    main()
