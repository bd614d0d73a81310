"""Host-side logic of bench.py that needs no GPU: the roofline summary picks the dominant KERNEL (k_pw_conv's timed roles
are merged into one family) and reports achieved algorithmic GB/s against the measured peak."""
import importlib.util
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("rubiks_bench", os.path.join(REPO, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_roofline_summary_merges_kernel_roles():
    bench = _bench()
    agg = {"pw_conv<dgrad>": {"bytes": 10 ** 10, "flops": 10 ** 12, "ms": 6.0, "launches": 98},
           "pw_conv<bn+relu>": {"bytes": 4 * 10 ** 9, "flops": 10 ** 11, "ms": 4.0, "launches": 47},
           "bn_backward": {"bytes": 23 * 10 ** 9, "flops": 0, "ms": 6.3, "launches": 94},
           "pw_conv_wgrad": {"bytes": 6 * 10 ** 9, "flops": 10 ** 11, "ms": 4.4, "launches": 59}}
    r = bench.summarize_roofline(agg)
    assert "k_pw_conv" in r["kernel"] and "k_pw3" in r["kernel"] and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["achieved"] - 1400.0) < 1e-6 and r["launches_per_step"] == 145
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert {k["kernel"] for k in r["all_kernels"]} == set(agg)
    assert abs(r["timed_ms_per_step"] - (6.0 + 4.0 + 6.3 + 4.4)) < 1e-6  # every call once (was double-counted)
    assert r["traffic"] is None or r["traffic"] > 0
    assert bench.summarize_roofline({}) is None
    r = bench.summarize_roofline({"bn_backward": agg["bn_backward"], "pw_conv": {"bytes": 1, "flops": 0, "ms": 0.1, "launches": 1}})
    assert r["kernel"] == "bn_backward" and r["traffic"] is None


def test_cli_defaults_and_multi_gpu_launch_mode():
    """No flags = N=1, a K/W that finish in minutes, fused SGD in both arms; --graph-multi auto resolves to the launch
    mode that was verified on that many GPUs (whole-step capture on <= 2, forward+backward capture beyond)."""
    bench = _bench()

    def parse(*argv):
        old, sys.argv = sys.argv, ["bench.py", *argv]
        try:
            return bench.parse()
        finally:
            sys.argv = old
    a = parse()
    assert (a.gpus, a.impl, a.batch, a.tier, a.variant, a.dtype, a.sgd) == (1, "ours", 32, "large", "rubiks3d", "bf16", "fused")
    assert a.warmup >= 3 and 1 <= a.steps <= 20
    assert a.graph_multi == "whole"
    assert parse("--gpus", "2").graph_multi == "whole"
    assert parse("--gpus", "4").graph_multi == "fwdbwd"
    assert parse("--gpus", "8").graph_multi == "fwdbwd"
    assert parse("--gpus", "8", "--graph-multi", "whole").graph_multi == "whole"
