#!/usr/bin/env bash
# Closing bench lines with torch's fused SGD in both arms (bench.py --sgd fused, the default from here on) and the A/B arm
# with the foreach kernel sequence of the earlier runs.  usage: gpu_r02_final5.sh [tag]
set -uo pipefail
T=${1:-r02zd}
O=gpurun_out; mkdir -p $O
run() { local name=$1; shift; timeout -k 10 400 python bench.py "$@" > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; python - $O/${T}_bench_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f" % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0)), d.get('dtype'), d.get('impl'), d['config'].get('optimizer'))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run c3_reference --impl reference --steps 8 --warmup 3
run c3_ours --steps 10 --warmup 3
run c3_ours_sgd_foreach --sgd foreach --steps 10 --warmup 3 --no-cpu-baseline
run c4_ours --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline
tail -2 $O/${T}_bench_c3_ours.err
