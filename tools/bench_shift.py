"""Op-level benchmark of the 3D shift: tiled sm_100a kernels vs generic gather kernels vs (if baseline/_ref
is present) the reference's own CUDA extension, on BASELINE config C1 and on every RubiksNet-Large
shift-layer geometry at a given clip batch.  Reports time per call and achieved algorithmic HBM GB/s
(fwd = (in+out) bytes, bwd = (x + og + x_grad) bytes).

    python tools/bench_shift.py [--batch 32] [--dtype float32|bfloat16] [--iters 20] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from rubiksnet_b200 import _lib  # noqa: E402
from rubiksnet_b200.shiftlib.rubiks3d.primitive import rubiks_shift_3d_backward, rubiks_shift_3d_forward  # noqa: E402

LARGE_LAYERS = [  # (name, C, H_in, spatial stride, number of such layers in RubiksNet-Large)
    ("layer0", 72, 112, 1, 1), ("layer1.0", 72, 112, 2, 1), ("layer1.x", 72, 56, 1, 2),
    ("layer2.0", 144, 56, 2, 1), ("layer2.x", 144, 28, 1, 7), ("layer3.0", 288, 28, 2, 1),
    ("layer3.x", 288, 14, 1, 35), ("layer4.0", 576, 14, 2, 1), ("layer4.x", 576, 7, 1, 2),
]


def reference_ext():
    ref = os.path.join(REPO, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "rubiksnet")):
        return None
    sys.path.insert(0, ref)
    try:
        import rubiksnet_cuda
        return rubiksnet_cuda
    except Exception as e:  # noqa: BLE001
        print("reference extension unavailable:", e, file=sys.stderr)
        return None


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
    times = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2]


def bench_case(name, N, T, C, H, S, dtype, iters, ref_ext, results):
    x = torch.randn(N, T, C, H, H, device="cuda").to(dtype)
    shift = torch.rand(3, C, device="cuda") * 2 - 1
    stride = (1, S, S)
    out = rubiks_shift_3d_forward(x, shift, stride, 0)
    og = torch.randn_like(out)
    es = x.element_size()
    fwd_bytes = (x.numel() + out.numel()) * es
    bwd_bytes = (2 * x.numel() + out.numel()) * es
    flush = fwd_bytes < (400 << 20)
    row = {"case": name, "N": N, "C": C, "H": H, "stride": S, "dtype": str(dtype).split(".")[-1],
           "fwd_MB": fwd_bytes / 1e6, "bwd_MB": bwd_bytes / 1e6, "l2_flush": flush}
    gin, gs = torch.empty_like(x), torch.empty_like(shift)
    for impl_name, impl in (("tiled", _lib.RB_IMPL_AUTO), ("generic", _lib.RB_IMPL_GENERIC)):
        _lib.set_impl(impl)
        tf = timeit(lambda: rubiks_shift_3d_forward(x, shift, stride, 0, output=out), iters, flush)
        tb = timeit(lambda: rubiks_shift_3d_backward(og, x, shift, stride, 0, True, x_grad_output=gin,
                                                     shift_grad_output=gs), iters, flush)
        row[impl_name] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_GBs": fwd_bytes / tf / 1e6, "bwd_GBs": bwd_bytes / tb / 1e6}
    _lib.set_impl(_lib.RB_IMPL_AUTO)
    if ref_ext is not None and dtype == torch.float32:
        r_out, r_gin, r_gs = torch.zeros_like(out), torch.zeros_like(x), torch.zeros_like(shift)

        def ref_fwd():  # the reference's Python wrapper allocates a zero-filled output every call (rubiksnet/utils.py:26)
            o = x.new_zeros(out.shape)
            ref_ext.rubiks_shift_3d_forward_float(x, shift, list(stride), [0, 0, 0], False, o)

        def ref_bwd():
            gi, gsh = x.new_zeros(x.shape), shift.new_zeros(shift.shape)
            ref_ext.rubiks_shift_3d_backward_float(x, shift, og, list(stride), [0, 0, 0], gi, gsh, True, 1.0, False)

        tf, tb = timeit(ref_fwd, max(3, iters // 4), flush), timeit(ref_bwd, max(3, iters // 4), flush)
        row["reference_ext"] = {"fwd_ms": tf, "bwd_ms": tb, "fwd_GBs": fwd_bytes / tf / 1e6, "bwd_GBs": bwd_bytes / tb / 1e6}
    results.append(row)
    msg = "%-10s N=%-3d C=%-3d H=%-3d s=%d %-8s" % (name, N, C, H, S, row["dtype"])
    for k in ("tiled", "generic", "reference_ext"):
        if k in row:
            msg += " | %s fwd %.3f ms %.0f GB/s bwd %.3f ms %.0f GB/s" % (k, row[k]["fwd_ms"], row[k]["fwd_GBs"],
                                                                       row[k]["bwd_ms"], row[k]["bwd_GBs"])
    print(msg, flush=True)
    del x, out, og, gin


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--dtype", default="both")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--no-ref", action="store_true")
    args = ap.parse_args()
    ref_ext = None if args.no_ref else reference_ext()
    dtypes = {"float32": [torch.float32], "bfloat16": [torch.bfloat16], "both": [torch.float32, torch.bfloat16]}[args.dtype]
    results = []
    print("device:", torch.cuda.get_device_name(0), "| reference ext:", ref_ext is not None)
    for dt in dtypes:
        bench_case("C1", 2, 8, 64, 56, 1, dt, args.iters, ref_ext, results)
        for name, C, H, S, _ in LARGE_LAYERS:
            bench_case(name, args.batch, 8, C, H, S, dt, args.iters, ref_ext, results)
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(results, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
