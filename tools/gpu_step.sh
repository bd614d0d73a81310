#!/usr/bin/env bash
# full GPU test suite + C3 / C4 bench lines (no CPU baseline leg) with the per-kernel table; TAG names the output files
set -uo pipefail
TAG=${1:-step}; O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/${TAG}_tests.log)"; grep -n "^FAILED\|^ERROR" $O/${TAG}_tests.log | head -20
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0), d.get('gpu_launches')))
r=d.get('roofline')
print("  top:", r['kernel'][:60], r['kernel_ms_per_step'], r['frac'])
for k in r['all_kernels'][:12]: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
}
for l in layer0 layer1.x layer2.x; do timeout -k 10 100 python tools/bench_pw.py --iters 10 --modes fwd,bn,res,dgrad,wgrad,wgrad_bn,cublas --only $l 2>&1 | grep layer; done | tee $O/${TAG}_pw.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err; summ $O/${TAG}_bench_c3.json; tail -2 $O/${TAG}_bench_c3.err | cut -c1-300
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; summ $O/${TAG}_bench_c4.json
