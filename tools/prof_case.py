"""One shift-layer geometry, a few forward + backward calls: the command to wrap in `ncu --set full -k regex:...`.

    python tools/prof_case.py --C 288 --H 14 --stride 1 --batch 32 --dtype bfloat16 --iters 3
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_backward, rubiks_shift_3d_forward  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=288)
    ap.add_argument("--H", type=int, default=14)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--dtype", default="bfloat16")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dt = getattr(torch, a.dtype)
    torch.manual_seed(0)
    x = torch.randn(a.batch, 8, a.C, a.H, a.H, device="cuda").to(dt)
    shift = torch.rand(3, a.C, device="cuda") * 2 - 1
    stride = (1, a.stride, a.stride)
    out = rubiks_shift_3d_forward(x, shift, stride, 0)
    og = torch.randn_like(out)
    gin, gs = torch.empty_like(x), torch.empty_like(shift)
    for _ in range(a.iters):
        rubiks_shift_3d_forward(x, shift, stride, 0, output=out)
        rubiks_shift_3d_backward(og, x, shift, stride, 0, True, x_grad_output=gin, shift_grad_output=gs)
    torch.cuda.synchronize()
    print("ok", tuple(x.shape), a.dtype)


if __name__ == "__main__":
    main()
