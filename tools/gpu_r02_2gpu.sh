#!/usr/bin/env bash
# 2 x B200 (gpurun --gpus 2): ours (whole-step graph incl. the NCCL all-reduce) and, with ARMS="reference ours", the reference
# under DDP on the same box.  usage: gpu_r02_2gpu.sh [tag] ; ARMS defaults to "ours"
set -uo pipefail
T=${1:-r02zc}; ARMS=${ARMS:-ours}
O=gpurun_out; mkdir -p $O
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for arm in $ARMS; do
  extra=""; [ $arm = reference ] && extra="--impl reference"
  timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM % 10)) bench.py --gpus 2 --steps 10 --warmup 3 $extra > $O/${T}_bench_c5_2gpu_$arm.json 2> $O/${T}_bench_c5_2gpu_$arm.err; echo "2gpu $arm exit=$?"
  python - $O/${T}_bench_c5_2gpu_$arm.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value'), d['config'].get('launch'))
except Exception as e: print("unreadable", e)
PY
done
tail -3 $O/${T}_bench_c5_2gpu_ours.err
