#!/usr/bin/env bash
# 2 x B200 (gpurun --gpus 2): ours (whole-step graph incl. the NCCL all-reduce) and the reference under DDP, same box
set -uo pipefail
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for arm in reference ours; do
  extra=""; [ $arm = reference ] && extra="--impl reference"
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM % 10)) bench.py --gpus 2 --steps 8 --warmup 3 $extra > $O/r02z_bench_2gpu_$arm.json 2> $O/r02z_bench_2gpu_$arm.err; echo "2gpu $arm exit=$?"
  python - $O/r02z_bench_2gpu_$arm.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config'].get('launch'))
except Exception as e: print("unreadable", e)
PY
done
tail -3 $O/r02z_bench_2gpu_ours.err
