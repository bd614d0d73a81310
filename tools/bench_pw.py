"""Op-level benchmark of the tcgen05 pointwise-conv kernels on every RubiksNet-Large block geometry: forward (plain /
bn+relu producer / fused 3D shift), input gradient and weight gradient, against torch.matmul (cuBLAS) on the same
NCHW bf16 tensors.  Reports time per call and achieved algorithmic HBM GB/s.

    python tools/bench_pw.py [--batch 32] [--iters 20] [--only layer3.x] [--modes fwd,shift,wgrad,...]
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import _lib, ops  # noqa: E402

LAYERS = [("layer0", 72, 112), ("layer1.x", 72, 56), ("layer2.x", 144, 28), ("layer3.x", 288, 14), ("layer4.x", 576, 7)]
BF = torch.bfloat16
_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
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
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default=None)
    ap.add_argument("--modes", default="fwd,fwd2,res,res2,bn,bn2,dgrad,dgrad2,shift,wgrad,wgrad_shift,cublas")
    ap.add_argument("--splits", type=int, default=0, help="rb_pw_conv_set_tuning: minimum output-channel splits (0 = auto)")
    ap.add_argument("--opstages", type=int, default=0, help="image kernel: operand ring depth (0 = auto)")
    ap.add_argument("--kc", type=int, default=0, help="image kernel: channels per K chunk, 16 or 32 (0 = auto)")
    ap.add_argument("--waitns", type=int, default=0, help="image kernel: mbarrier suspend-time hint in ns (0 = default)")
    ap.add_argument("--geom", default="", help="extra geometry name,C,H,W (e.g. pad14,288,25,8 = a 14x14 map at a 200-element pitch)")
    ap.add_argument("--wg", default="", help="tensor-map weight gradient knobs: burst,l2_256,max_stages")
    ap.add_argument("--no-tma", action="store_true", help="tensor-map TMA schedules (k_pw3 / k_wg3) off: first kernels everywhere")
    a = ap.parse_args()
    if a.no_tma:
        _lib.set_pw_tma(False)
    if a.wg:
        _lib.lib().rb_pw_conv_wgrad_set_tuning(*[int(v) for v in a.wg.split(",")])
    _lib.lib().rb_pw_conv_set_tuning(a.splits)
    _lib.lib().rb_pw_conv_image_set_tuning(a.opstages, a.kc, a.waitns)
    modes = a.modes.split(",")
    T = 8
    print("device:", torch.cuda.get_device_name(0))
    layers = [(n, c, h, h) for n, c, h in LAYERS]
    if a.geom:
        gn, gc, gh, gw = a.geom.split(",")
        layers.append((gn, int(gc), int(gh), int(gw)))
    for name, c, h, wd in layers:
        if a.only and a.only != name:
            continue
        ni = a.batch * T
        x = torch.randn(ni, c, h, wd, device="cuda").to(BF)
        w = torch.randn(c, c, device="cuda") / c ** 0.5
        res = torch.randn(ni, c, h, wd, device="cuda").to(BF)
        g = torch.randn(ni, c, h, wd, device="cuda").to(BF)
        shift = torch.rand(3, c, device="cuda") * 2 - 1
        sb = torch.stack([torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")], dim=1).contiguous()
        unit = x.numel() * 2
        flush = unit * 3 < (300 << 20)
        w_nk, w_kn = ops.pw_weight_pack(w)  # what the block path feeds the kernels (bf16, both orientations)
        cases = {
            "fwd": (lambda: ops.pw_conv(x, w_nk), 2 * unit),
            "res": (lambda: ops.pw_conv(x, w_nk, residual=res), 3 * unit),
            "bn": (lambda: ops.pw_conv(x, w_nk, in_scale_bias=sb), 2 * unit),
            "res_st": (lambda: ops.pw_conv(x, w_nk, residual=res, stats=True), 3 * unit),  # + BatchNorm statistics of the output
            "bn_st": (lambda: ops.pw_conv(x, w_nk, in_scale_bias=sb, stats=True), 2 * unit),
            "shift": (lambda: ops.shift3d_pw_conv(x, shift, w_nk, res, T), 3 * unit),
            "dgrad": (lambda: ops.pw_conv(g, w_kn), 2 * unit),
            "wgrad": (lambda: ops.pw_conv_wgrad(g, x), 2 * unit),
            "wgrad_bn": (lambda: ops.pw_conv_wgrad(g, x, in_scale_bias=sb), 2 * unit),
            "wgrad_shift": (lambda: ops.shift3d_pw_conv_wgrad(g, x, shift, T), 2 * unit),
        }
        if ops.pw_image_supported(ni, c, c, h * wd, True):
            img_f, img_b = ops.pw_weight_images(w)
            cases.update({
                "fwd2": (lambda: ops.pw_conv(x, img_f), 2 * unit),
                "res2": (lambda: ops.pw_conv(x, img_f, residual=res), 3 * unit),
                "bn2": (lambda: ops.pw_conv(x, img_f, in_scale_bias=sb), 2 * unit),
                "dgrad2": (lambda: ops.pw_conv(g, img_b), 2 * unit),
            })
        wb = w.to(BF)
        xb = x.view(ni, c, -1)
        cases["cublas"] = (lambda: torch.matmul(wb, xb), 2 * unit)
        msg = "%-9s C=%-3d H=%-3d unit %.1f MB |" % (name, c, h, unit / 1e6)
        for m in modes:
            if m not in cases:
                continue
            fn, nbytes = cases[m]
            ms = timeit(fn, a.iters, flush)
            msg += " %s %.3f ms %.0f GB/s |" % (m, ms, nbytes / ms / 1e6)
        print(msg, flush=True)


if __name__ == "__main__":
    main()
