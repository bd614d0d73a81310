#!/usr/bin/env python
"""Headline benchmark: clips/s of one training step (forward + backward + SGD update) of RubiksNet-Large,
8 frames x 224^2, bf16 activations, synthetic clips, random-init weights -- BASELINE.json config C3 (C5 when
--gpus > 1: batch data parallel, one process per GPU, NCCL gradient all-reduce).

    python bench.py --gpus 1 --steps K --warmup W                 # this repo (rubiksnet_b200)
    python bench.py --impl reference --gpus N --steps K --warmup W  # the unmodified reference (baseline/_ref)
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value        clips/s, whole job, inputs already resident in HBM (device-timed, max over ranks)
  e2e          same step driven from pinned HOST clips: H2D copy of every batch and a D2H read of the
               loss inside the timed region
  roofline     achieved algorithmic HBM GB/s of the dominant hand-written kernel, measured live with CUDA
               events around every librubiks_b200 launch on the launching stream (all kernels listed)
  cpu_baseline the CPU port (oracle/ C restatement of the shift + PyTorch CPU ops for the rest of the
               network) timed on this box's host cores on a bounded sample
  gpu_launches number of librubiks_b200 kernels launched inside the timed region
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

NUM_CLASSES = 174  # Something-Something-v2, as in the reference's checkpoints
FRAMES = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU (weak scaling)")
    ap.add_argument("--tier", default="large")
    ap.add_argument("--variant", default="rubiks3d")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    ap.add_argument("--cpu-sample-clips", type=int, default=2, help="clips per CPU-baseline pass")
    ap.add_argument("--cpu-sample-reps", type=int, default=10, help="passes of the CPU baseline (bounded sample: ~10-20 s)")
    ap.add_argument("--infer", action="store_true",
                    help="inference: eval-mode forward under no_grad (BASELINE C2 with --tier tiny --dtype fp32)")
    ap.add_argument("--ref-autocast", action="store_true",
                    help="--impl reference only: run the reference model under bf16 autocast with .float() casts around its "
                         "(fp32-only) shift ops -- the like-for-like precision arm of SURVEY 8d(1)")
    ap.add_argument("--graph-multi", default="auto", choices=["auto", "whole", "fwdbwd"],
                    help="N > 1: capture the whole step including the NCCL all-reduce and the SGD update (whole), or only "
                         "forward + backward with the exchange and the update eager (fwdbwd, round-1 behaviour).  auto = whole "
                         "on 2 GPUs, fwdbwd beyond: the whole-step capture was only ever run on 2 GPUs (2 023 vs 2 013 clips/s "
                         "there, profiles/r02zd_bench_c5_2gpu_ours.json / r02ze_bench_c5_2gpu_fwdbwd.json), the fwdbwd path on "
                         "2 and 8 (SCALE_r01.json)")
    ap.add_argument("--sgd", default="fused", choices=["fused", "foreach"],
                    help="torch.optim.SGD implementation of both arms: one fused multi-tensor kernel per chunk (default) or the "
                         "foreach kernel sequence (the setting of the runs before r02zd)")
    ap.add_argument("--graph", default="on", choices=["on", "off"],
                    help="replay the whole step from a CUDA graph (rubiksnet_b200.graph.GraphedStep); off = eager launches")
    ap.add_argument("--pdl", default="on", choices=["on", "off"],
                    help="programmatic dependent launch of the library's kernels (rb_set_dependent_launch); off = A/B arm")
    ap.add_argument("--tma", default="on", choices=["on", "off"],
                    help="tensor-map TMA schedules of the 1x1 convs / weight gradients on 16-byte-pitch maps (k_pw3, k_wg3); off = A/B arm")
    args = ap.parse_args()
    if args.graph_multi == "auto":
        args.graph_multi = "whole" if args.gpus <= 2 else "fwdbwd"
    return args


# --------------------------------------------------------------------------------------------- helpers

def dist_setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return world, rank, local


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, "/tmp/rubiks_bench_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [s.strip() for s in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_batch(batch, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    clips = torch.randn(batch, FRAMES, 3, 224, 224, generator=g)
    labels = torch.randint(0, NUM_CLASSES, (batch,), generator=g)
    return clips, labels


# ------------------------------------------------------------------------------------ training steps

class Trainer:
    """fwd + bwd + (all-reduce) + SGD on one GPU.  `kind` = "ours" (rubiksnet_b200) or "reference"."""

    def __init__(self, kind, args, world):
        import torch
        self.torch, self.kind, self.world, self.args = torch, kind, world, args
        torch.manual_seed(0)
        torch.backends.cudnn.benchmark = True  # as the reference does (scripts/test_models.py:46)
        if kind == "ours":
            import rubiksnet_b200 as rb
            from rubiksnet_b200.dp import FlatGradAllReduce
            self.net = rb.RubiksNet(tier=args.tier, num_classes=NUM_CLASSES, num_frames=FRAMES,
                                    variant=args.variant).cuda().train()
            self.reducer = None if args.infer else FlatGradAllReduce(self.net)
            self.fwd_net = self.net
            self.autocast = args.dtype == "bf16"
        else:
            sys.path.insert(0, os.path.join(REPO, "baseline", "_ref"))
            import contextlib
            import io
            with contextlib.redirect_stdout(io.StringIO()):  # the reference prints banners
                from rubiksnet.models import RubiksNet  # noqa: E402  (the reference package)
                self.net = RubiksNet(tier=args.tier, num_classes=NUM_CLASSES, num_frames=FRAMES,
                                     variant=args.variant).cuda().train()
            self.reducer = None
            self.fwd_net = self.net
            if world > 1 and not args.infer:
                self.fwd_net = torch.nn.parallel.DistributedDataParallel(self.net, device_ids=[torch.cuda.current_device()])
            # the reference's 3D shift is float/double only (primitive.py:66-75) and its 2D shift has no bf16 path: stock
            # = fp32 end to end.  --ref-autocast: bf16 autocast for the library ops with the shift ops computed in fp32
            # between casts (the package itself stays unmodified: the casts are nn.Module forward hooks on its shift modules)
            self.autocast = bool(args.ref_autocast) and args.dtype == "bf16"
            if self.autocast:
                def to_float(mod, inputs):
                    mod._rb_in_dtype = inputs[0].dtype
                    return (inputs[0].float(),) + tuple(inputs[1:])

                def to_input_dtype(mod, inputs, output):
                    return output.to(mod._rb_in_dtype)
                for m in self.net.modules():
                    if "shift" in m._parameters:  # RubiksShift3D / RubiksShift2D of the reference
                        m.register_forward_pre_hook(to_float)
                        m.register_forward_hook(to_input_dtype)
        if args.infer:
            self.net.eval()
            self.opt = None
            self.loss_fn = None
            self.graphed = None
            return
        shift_params = [p for n, p in self.net.named_parameters() if n.endswith("shift")]
        other = [p for n, p in self.net.named_parameters() if not n.endswith("shift")]
        # shift parameters get lr * 0.01 as in scripts/example_finetune.py:49-64
        self.opt = torch.optim.SGD([{"params": shift_params, "lr": 1e-4}, {"params": other}], lr=1e-2,
                                   momentum=0.9, weight_decay=1e-4, **({"fused": True} if args.sgd == "fused" else {"foreach": True}))
        self.loss_fn = torch.nn.CrossEntropyLoss()
        self.graphed = None

    def runner(self, clips, labels):
        """The callable the timed loops drive: the CUDA-graph replay of `step` (ours, --graph on) or `step` itself."""
        # N > 1: eager launches -- the NCCL all-reduce stays outside any capture (a whole-step capture that includes the
        # collective hung on 2 GPUs, gpurun_out/s28)
        if self.kind != "ours" or self.args.graph != "on":
            return self.step
        from rubiksnet_b200.graph import GraphedStep
        if self.world == 1 or self.args.infer or self.args.graph_multi == "whole":
            if self.graphed is None:
                self.graphed = GraphedStep(self.step, clips, labels, warmup=2)
            return self.graphed
        # N > 1: forward + backward replayed from a graph; the gradient all-reduce and the SGD update stay eager (a
        # whole-step capture including the NCCL collective ran, but hung at process teardown: gpurun_out/s31)
        if self.graphed is None:
            self.graphed = GraphedStep(self.fwd_bwd, clips, labels, warmup=2)
            self.static_grads = [p.grad for p in self.reducer.params]  # where the captured backward writes

        def run(c, l):
            loss = self.graphed(c, l)
            for p, g in zip(self.reducer.params, self.static_grads):
                p.grad = g
            self.reducer.all_reduce()
            self.opt.step()
            return loss
        return run

    def close(self):
        """Releases the captured graph BEFORE the process group goes away: a CUDA graph that holds NCCL kernels keeps the
        communicator busy, and destroying the communicator first hung the process at exit in round 1 (gpurun_out/s31)."""
        torch = self.torch
        torch.cuda.synchronize()
        if self.graphed is not None:
            g, self.graphed = self.graphed, None
            g.graph.reset()
            del g
        torch.cuda.synchronize()

    def infer(self, clips, labels):
        torch = self.torch
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            logits = self.fwd_net(clips)
        return logits.float().logsumexp(1).mean()  # a scalar the e2e loop reads back (stands in for the loss)

    def fwd_bwd(self, clips, labels):
        torch = self.torch
        if self.reducer is not None:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=False)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            logits = self.fwd_net(clips)
        loss = self.loss_fn(logits.float(), labels)
        loss.backward()
        return loss

    def step(self, clips, labels):
        if self.args.infer:
            return self.infer(clips, labels)
        loss = self.fwd_bwd(clips, labels)
        if self.reducer is not None:
            self.reducer.all_reduce()
        self.opt.step()
        return loss


def time_resident(tr, args, world, clips, labels):
    import torch
    run = tr.runner(clips, labels)
    for _ in range(args.warmup):
        run(clips, labels)
    barrier(world)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        run(clips, labels)
    b.record()
    barrier(world)
    torch.cuda.synchronize()
    return max_over_ranks(a.elapsed_time(b), world)


def time_e2e(tr, args, world, host_batches):
    """Pinned host clips -> H2D on a copy stream (double-buffered) -> step -> loss.item() every step."""
    import torch
    copy_stream = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def upload(i):
        hc, hl = host_batches[i % len(host_batches)]
        with torch.cuda.stream(copy_stream):
            dc = hc.cuda(non_blocking=True)
            dl = hl.cuda(non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return dc, dl, ev

    step = tr.runner(*upload(0)[:2])

    def run(n_steps):
        nxt = upload(0)
        total = 0.0
        for i in range(n_steps):
            dc, dl, ev = nxt
            main.wait_event(ev)
            dc.record_stream(main)
            dl.record_stream(main)
            if i + 1 < n_steps:
                nxt = upload(i + 1)  # overlaps with this step's compute
            total += step(dc, dl).item()  # D2H read of the step's result
        return total

    run(max(1, min(args.warmup, 2)))
    barrier(world)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.steps)
    b.record()
    barrier(world)
    torch.cuda.synchronize()
    hc, hl = host_batches[0]
    return max_over_ranks(a.elapsed_time(b), world), hc.numel() * hc.element_size() + hl.numel() * hl.element_size()


def kernel_roofline(tr, clips, labels):
    """Brackets every librubiks_b200 call of ONE extra step with CUDA events on the launching stream
    (rubiksnet_b200._lib.timing) and reports the dominant hand-written kernel (by total time) against the measured
    HBM copy peak: achieved = algorithmic bytes (each operand tensor once) / event time."""
    import torch
    from rubiksnet_b200 import _lib
    _lib.timing.start()
    try:
        tr.step(clips, labels)
        torch.cuda.synchronize()
    finally:
        agg = _lib.timing.stop()
    return summarize_roofline(agg)


def summarize_roofline(agg):
    """agg: {call name: {bytes, flops, ms, launches}} of one step -> the `roofline` object of the JSON line."""
    if not agg:
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback B200_PROFILING.md")
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    # the dominant KERNEL: the timed call names split k_pw_conv / k_pw_wgrad by role (dgrad, +residual, bn+relu producer)
    def family(name):
        if name.startswith("pw_conv_tf32"):
            return "pw_conv_tf32"
        return "pw_conv_wgrad" if name.startswith("pw_conv_wgrad") else "pw_conv" if name.startswith("pw_conv") else name
    fam = {}
    for name, d in agg.items():
        f = fam.setdefault(family(name), {"bytes": 0, "flops": 0, "ms": 0.0, "launches": 0})
        for key in f:
            f[key] += d[key]
    top_family = max(fam, key=lambda k: fam[k]["ms"])
    if top_family in ("pw_conv", "pw_conv_wgrad", "pw_conv_tf32"):
        # the bf16 families run on two schedules each: register-staged (k_pw_conv / k_pw_wgrad; k_pw2 on 7x7 maps) and, on
        # maps with a 16-byte row pitch, tensor-map TMA (k_pw3 / k_wg3)
        label = {"pw_conv": "1x1-conv forward / input-gradient GEMMs k_pw_conv + k_pw3 + k_pw2 (all roles: ",
                 "pw_conv_wgrad": "1x1-conv weight-gradient GEMMs k_pw_wgrad + k_wg3 (all roles: ",
                 "pw_conv_tf32": "k_pw_tf32 (all roles: "}[top_family]
        label += ", ".join(sorted(n for n in agg if family(n) == top_family)) + ")"
        agg = dict(agg)
        agg[label] = fam[top_family]
        top = label
    else:
        top = top_family

    def entry(name):
        d = agg[name]
        gbs = d["bytes"] / d["ms"] / 1e6
        e = {"kernel": name, "achieved": round(gbs, 1), "frac": round(gbs / peak, 4), "launches_per_step": d["launches"],
             "algorithmic_bytes_per_step": d["bytes"], "kernel_ms_per_step": round(d["ms"], 3)}
        if d["flops"]:
            e["tensor_tflops"] = round(d["flops"] / d["ms"] / 1e9, 1)
            # TF32 runs at half the bf16 rate: the measured bf16 peak / 2 is the denominator for the tf32 kernel
            e["tensor_frac"] = round(d["flops"] / d["ms"] / 1e9 / (tf_peak / 2 if "tf32" in name else tf_peak), 4)
        return e

    out = {"bound": "hbm", "peak": peak, "unit": "GB/s", "traffic": None, "peak_source": peak_src}
    out.update(entry(top))
    try:  # DRAM bytes of one launch of that kernel family from the committed ncu --set full capture
        tf = [f for f in sorted(os.listdir(os.path.join(REPO, "profiles"))) if f.endswith("_traffic.json")][-1]
        t = json.load(open(os.path.join(REPO, "profiles", tf))).get(top_family)
        if t:
            out["traffic"] = t["dram_bytes_per_launch"]
            out["traffic_detail"] = dict(t, source="profiles/" + tf)
    except Exception:  # noqa: BLE001
        pass
    out["all_kernels"] = [entry(k) for k in sorted(agg, key=lambda k: -agg[k]["ms"]) if k != top]
    # sum over the individual calls only (the merged family entry added above is not a second kernel); these are
    # event-bracketed EAGER launches, so the sum exceeds the graph-replayed step time
    out["timed_ms_per_step"] = round(sum(d["ms"] for d in fam.values()), 3)
    out["timed_note"] = "CUDA-event time of every librubiks_b200 call in ONE extra eager step (not the graph replay)"
    return out


# ------------------------------------------------------------------------------------ CPU port arm

def cpu_port_clips_per_s(args, clips_in_sample):
    """RubiksNet on the host cores: PyTorch CPU ops for conv/BN/ReLU and the C oracle (OpenMP, all cores) for
    the shift, which the reference cannot run on CPU at all (primitive.py:61).  One fwd+bwd of a bounded
    sample."""
    import torch
    import oracle
    import rubiksnet_b200 as rb
    import rubiksnet_b200.shiftlib.rubiks3d.layer as layer3d
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle.set_num_threads(cores)

    class OracleShift3D(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, shift, stride, padding, normalize_grad, t_factor, quantize):
            ctx.cfg = (stride, padding, normalize_grad, t_factor, quantize)
            ctx.save_for_backward(x, shift)
            return torch.from_numpy(oracle.shift3d_forward(x, shift, stride, padding, quantize))

        @staticmethod
        def backward(ctx, g):
            x, shift = ctx.saved_tensors
            stride, padding, ng, tf, q = ctx.cfg
            gin, gs = oracle.shift3d_backward(x, shift, g.contiguous(), stride, padding, ng, tf, q)
            return torch.from_numpy(gin), torch.from_numpy(gs), None, None, None, None, None

    def cpu_shift(x, shift, stride=1, padding=0, normalize_grad=True, normalize_t_factor=1.0, quantize=False):
        return OracleShift3D.apply(x.contiguous(), shift, stride, padding, normalize_grad, normalize_t_factor, quantize)

    torch.manual_seed(0)
    net = rb.RubiksNet(tier=args.tier, num_classes=NUM_CLASSES, num_frames=FRAMES, variant="rubiks3d").train()
    for m in net.modules():
        if isinstance(m, layer3d.RubiksShiftBase):
            m.shift_function = cpu_shift
    clips, labels = synthetic_batch(clips_in_sample, 1)
    reps = max(1, int(getattr(args, "cpu_sample_reps", 1)))
    t0 = time.perf_counter()
    for _ in range(reps):
        net.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(net(clips), labels)
        loss.backward()
    dt = time.perf_counter() - t0
    return clips_in_sample * reps / dt, cores, dt


# --------------------------------------------------------------------------------------------- main

def main():
    args = parse()
    import torch
    have_cuda = torch.cuda.is_available()
    metric = "clips/sec (%s) RubiksNet-%s 8x224^2" % ("inference" if args.infer else "fwd+bwd", args.tier.capitalize())
    if args.infer:
        cfg_name = "C2" if (args.tier, args.variant, args.dtype) == ("tiny", "rubiks3d", "fp32") else "C2-like"
        what = "eval-mode forward under no_grad"
    else:
        cfg_name = {"rubiks3d": "C3" if args.gpus == 1 else "C5", "rubiks3d-aq": "C4"}.get(args.variant, "C3")
        if args.tier != "large" or args.dtype != "bf16":
            cfg_name += "-like"
        what = "fwd+bwd+SGD"
    config = {"workload": "BASELINE %s: RubiksNet-%s %s, 8 frames x 224^2, %s, %s, synthetic clips, random init"
                          % (cfg_name, args.tier.capitalize(), args.variant, args.dtype, what),
              "clips_per_gpu": args.batch, "frames": FRAMES, "num_classes": NUM_CLASSES,
              "parallelism": "dp%d (batch sharded, NCCL grad all-reduce)" % args.gpus,
              "launch": ("eager" if args.graph != "on" or args.impl != "ours" else
                         "whole step replayed from one CUDA graph" if args.gpus == 1 or args.graph_multi == "whole" else
                         "forward + backward replayed from one CUDA graph, NCCL all-reduce + SGD eager"),
              "l2": "working set per step (>10 GB of activations) far exceeds the 126 MB L2; no flush needed"}
    if not args.infer:
        config["optimizer"] = "torch.optim.SGD(momentum=0.9, weight_decay=1e-4, %s=True), both arms" % args.sgd

    if args.impl == "reference":
        ref_ok = os.path.isdir(os.path.join(REPO, "baseline", "_ref", "rubiksnet"))
        use_cuda = have_cuda and ref_ok and args.ref_device != "cpu"
        if not use_cuda:
            # fallback: CPU port (oracle) -- rank 0 only
            if int(os.environ.get("RANK", "0")) != 0:
                return
            v, cores, dt = cpu_port_clips_per_s(args, args.cpu_sample_clips)
            line = {"impl": "reference", "metric": metric, "value": v, "unit": "clips/s", "n_gpus": args.gpus,
                    "steps": 1, "warmup": 0, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                    "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                                     "sample": "%d x %d clip(s) fwd+bwd, oracle C shift + PyTorch CPU ops" % (args.cpu_sample_reps, args.cpu_sample_clips)},
                    "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
            print(json.dumps(line))
            return

    world, rank, local = dist_setup(args)
    assert have_cuda, "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    if args.impl == "ours":
        from rubiksnet_b200 import _lib as _rb_lib
        _rb_lib.set_dependent_launch(args.pdl == "on")
        config["dependent_launch"] = args.pdl
        if args.tma == "off":
            _rb_lib.set_pw_tma(False)
            config["tensor_map_schedules"] = "off"
    tr = Trainer("ours" if args.impl == "ours" else "reference", args, world)
    clips, labels = synthetic_batch(args.batch, 100 + rank)
    dclips, dlabels = clips.cuda(), labels.cuda()

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = 0
    if args.impl == "ours":
        from rubiksnet_b200 import _lib
    # ---- resident-input timing
    for _ in range(1):
        tr.step(dclips, dlabels)  # cudnn autotune / allocator warm-up outside everything
    if sampler:
        sampler.start()
    if args.impl == "ours":
        # launches are counted over the timed steps only: reset after warm-up inside time_resident is not
        # possible without a hook, so count (warmup + steps) and scale
        _lib.reset_launch_count()
    ms = time_resident(tr, args, world, dclips, dlabels)
    if args.impl == "ours":
        if tr.graphed is not None:  # replays do not pass through the library's host-side counter
            launches0 = tr.graphed.launches_per_replay * args.steps
        else:
            launches0 = _lib.launch_count() * args.steps // (args.steps + args.warmup)
    clocks = sampler.stop() if sampler else None
    total_clips = args.batch * world * args.steps
    value = total_clips / (ms / 1e3)

    e2e = None
    if not args.no_e2e:
        host = []
        for i in range(2):
            c, l = synthetic_batch(args.batch, 200 + 10 * rank + i)
            host.append((c.pin_memory(), l.pin_memory()))
        ems, h2d = time_e2e(tr, args, world, host)
        e2e = {"value": total_clips / (ems / 1e3), "unit": "clips/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "ms_per_step": ems / args.steps}

    line = {"metric": metric, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if tr.autocast else "f32", "data": "synthetic", "config": config,
            "clocks": clocks, "e2e": e2e}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["config"]["reference"] = (
            "unmodified reference package + its CUDA extension (baseline/_ref, built for sm_100); " +
            ("bf16 autocast for the library ops, its shift ops in fp32 between casts (--ref-autocast, SURVEY 8d(1))"
             if tr.autocast else "stock path: fp32 end to end (its 3D shift has no 16-bit path)") + "; DDP for n_gpus>1")
        line["gpu_launches"] = None
    else:
        line["impl"] = "rubiksnet_b200"
        line["gpu_launches"] = launches0
        roof = kernel_roofline(tr, dclips, dlabels)  # every rank: the extra step contains the gradient all-reduce
        if rank == 0:
            line["roofline"] = roof
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            v, cores, dt = cpu_port_clips_per_s(args, args.cpu_sample_clips)
            line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                                    "sample": "%d x %d clip(s), fwd+bwd of the same network in fp32: oracle C shift (OpenMP) + "
                                              "PyTorch CPU conv/BN; %.1f s" % (args.cpu_sample_reps, args.cpu_sample_clips, dt)}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    tr.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
