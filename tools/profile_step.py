"""Kernel-level breakdown of one training step (torch.profiler, CUDA time per kernel name).
    python tools/profile_step.py [--impl ours|reference] [--batch 32] [--dtype bf16|fp32] [--top 40]
"""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--tier", default="large")
    ap.add_argument("--variant", default="rubiks3d")
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    args = argparse.Namespace(tier=a.tier, variant=a.variant, dtype=a.dtype, batch=a.batch)
    tr = bench.Trainer(a.impl, args, 1)
    clips, labels = bench.synthetic_batch(a.batch, 1)
    clips, labels = clips.cuda(), labels.cuda()
    for _ in range(3):
        tr.step(clips, labels)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.step(clips, labels)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name, [0.0, 0])
            r[0] += ev.device_time
            r[1] += 1
    total = sum(v[0] for v in rows.values())
    print("impl=%s batch=%d dtype=%s: total CUDA kernel time %.2f ms over %d kernel launches" %
          (a.impl, a.batch, a.dtype, total / 1e3, sum(v[1] for v in rows.values())))
    for name, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print("%8.3f ms %5.1f%% %5d  %s" % (t / 1e3, 100 * t / total, n, name[:150]))


if __name__ == "__main__":
    main()
