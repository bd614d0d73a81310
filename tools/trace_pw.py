"""Pipeline timeline of one k_pw_conv launch (debug): per-CTA globaltimer stamps of setup / weights / per-tile
producer-done, MMA-issued, epilogue start/end.   python tools/trace_pw.py [--C 288 --H 14 --batch 32 --mode fwd]"""
import argparse
import ctypes
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import _lib, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=288)
    ap.add_argument("--H", type=int, default=14)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--mode", default="fwd")
    ap.add_argument("--wbf16", action="store_true", help="bf16 weights packed by rb_pw_weight_pack")
    ap.add_argument("--dbg", type=int, default=0, help="debug library: work-skipping flags of the image kernel")
    ap.add_argument("--reps", type=int, default=0, help="also time `reps` launches with CUDA events (L2 flushed in between)")
    ap.add_argument("--opstages", type=int, default=0)
    ap.add_argument("--kc", type=int, default=0)
    ap.add_argument("--waitns", type=int, default=0)
    ap.add_argument("--v2", action="store_true", help="second-generation kernel (packed weight images, csrc/pw_conv2.cu)")
    a = ap.parse_args()
    # the stamps only exist in the debug library (python -m rubiksnet_b200.build --trace); point the loader at it
    _lib.LIB_PATH = _lib.LIB_PATH.replace(".so", "_trace.so")
    assert os.path.exists(_lib.LIB_PATH), "build the debug library first: python -m rubiksnet_b200.build --trace"
    L = _lib.lib()
    L.rb_debug_pw_trace.argtypes = [ctypes.c_void_p]
    L.rb_debug_pw_trace.restype = None
    ni = a.batch * 8
    x = torch.randn(ni, a.C, a.H, a.H, device="cuda").bfloat16()
    w = torch.randn(a.C, a.C, device="cuda") / a.C ** 0.5
    if a.wbf16:
        w, w_kn = ops.pw_weight_pack(w)
    sb = torch.stack([torch.rand(a.C, device="cuda") + 0.5, torch.randn(a.C, device="cuda")], dim=1).contiguous()
    if a.v2:
        L.rb_pw_conv_image_set_tuning.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.rb_pw_conv_image_set_tuning(a.opstages, a.kc, a.waitns)
        L.rb_debug_pw_flags.argtypes = [ctypes.c_int]
        L.rb_debug_pw_flags.restype = None
        L.rb_debug_pw_flags(a.dbg)
        img_f, img_b = ops.pw_weight_images(w)
        res = torch.randn_like(x)
        fn = {"fwd": lambda: ops.pw_conv(x, img_f), "bn": lambda: ops.pw_conv(x, img_f, in_scale_bias=sb),
              "res": lambda: ops.pw_conv(x, img_f, residual=res), "dgrad": lambda: ops.pw_conv(x, img_b)}[a.mode]
        for _ in range(3):
            fn()
        if a.reps:
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
            ts = []
            for _ in range(a.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            print("dbg=%d %s C=%d H=%d: median %.1f us" % (a.dbg, a.mode, a.C, a.H, ts[len(ts) // 2] * 1e3))
        trace = torch.zeros(148 * 128, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        L.rb_debug_pw_trace(ctypes.c_void_p(trace.data_ptr()))
        fn()
        torch.cuda.synchronize()
        L.rb_debug_pw_trace(None)
        t = trace.view(-1, 128).cpu()
        t = t[t[:, 0] > 0]
        t0 = int(t[:, 0].min())
        us = lambda v: (int(v) - t0) / 1e3 if int(v) else float("nan")  # noqa: E731
        print("v2 %s C=%d H=%d: CTAs traced %d, kernel span %.1f us" % (a.mode, a.C, a.H, t.shape[0], (int(t[:, 3].max()) - t0) / 1e3))
        names = ["mma-start", "stage0-full", "last-full", "epi-start", "staged", "stored", "tma-first", "tma-last", "rel-first-raw",
                 "rel-last-done", "res-ready"]
        for cta in (0, t.shape[0] // 2, t.shape[0] - 1):
            r = t[cta]
            print("CTA %d: start %.1f mma-ready %.1f weights %.1f end %.1f" % (cta, us(r[0]), us(r[1]), us(r[2]), us(r[3])))
            for it in range(4):
                b = 8 + it * 12
                if int(r[b]) == 0:
                    break
                print("   tile %d: " % it + "  ".join("%s %.1f" % (n, us(r[b + i])) for i, n in enumerate(names)))
            print("   tile 2 TMA warp per stage (before / after its raw_empty wait): " +
                  "  ".join("%.2f/%.2f" % (us(r[56 + 2 * st]), us(r[57 + 2 * st])) for st in range(9) if int(r[56 + 2 * st])))
            print("   tile 2 stage 3: arrival of the 8 relayout warps: " +
                  "  ".join("%.2f" % us(r[74 + w] if w < 6 else r[98 + w - 6]) for w in range(8)))
            print("   tile 2 relayout warp 0: the other wait seen at: " +
                  "  ".join("st%d %.2f" % (st, us(r[80 + 2 * st])) for st in range(12) if int(r[80 + 2 * st])))
            print("   tile 2 relayout warp 0, its stages (raw_full seen / arrived): " +
                  "  ".join("st%d %.2f/%.2f" % (st, us(r[104 + 2 * st]), us(r[105 + 2 * st])) for st in range(12) if int(r[104 + 2 * st])))
        return
    fn = {"fwd": lambda: ops.pw_conv(x, w), "bn": lambda: ops.pw_conv(x, w, in_scale_bias=sb),
          "dgrad": (lambda: ops.pw_conv(x, w_kn)) if a.wbf16 else (lambda: ops.pw_conv(x, w, transposed=True))}[a.mode]
    for _ in range(3):
        fn()
    trace = torch.zeros(148 * 4 * 128, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    L.rb_debug_pw_trace(ctypes.c_void_p(trace.data_ptr()))
    fn()
    torch.cuda.synchronize()
    L.rb_debug_pw_trace(None)
    t = trace.view(-1, 128).cpu()
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    names = {0: "start", 1: "setup-done(prod)", 2: "weights-staged", 3: "end"}
    print("CTAs traced:", t.shape[0], " kernel span %.1f us" % ((int(t[:, 3].max()) - t0) / 1e3))
    for cta in (0, 1, t.shape[0] // 2, t.shape[0] - 1):
        r = t[cta]
        print("CTA %d:" % cta, " ".join("%s=%.1f" % (names[i], (int(r[i]) - t0) / 1e3) for i in (0, 1, 2, 3)))
        for it in range(7):
            ev = [int(r[4 + it * 8 + j]) for j in range(8)]
            if not any(ev):
                break
            f = lambda v: "%.1f" % ((v - t0) / 1e3) if v else "-"
            print("   tile %d: mma-start %s  stage0-full %s  last-stage-full %s  mma-issued %s | producer(w0)-done %s | epi-start %s  epi-end %s" % (
                it, f(ev[4]), f(ev[5]), f(ev[6]), f(ev[0]), f(ev[3]), f(ev[1]), f(ev[2])))
        if cta == 0:  # per-stage handoffs of tile 2
            for st in range(8):
                ev = [int(r[64 + st * 8 + j]) for j in range(6)]
                if not any(ev):
                    break
                f = lambda v: "%.2f" % ((v - t0) / 1e3) if v else "-"
                print("      tile 2 stage %d: mma full-seen %s  issued+commit %s | producer: enter %s  landed %s  arrived %s  next-issued %s" % (
                    st, f(ev[0]), f(ev[1]), f(ev[2]), f(ev[3]), f(ev[4]), f(ev[5])))


if __name__ == "__main__":
    main()
