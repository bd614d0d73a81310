#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02ad_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02ad_tests.log)"; grep -n "^FAILED\|Error" $O/r02ad_tests.log | head -20
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ad_bench_c4.json 2> $O/r02ad_bench_c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ad_bench_c4.json').read().strip().splitlines()[-1])
print("C4 value %.1f ms/step %.3f e2e %.1f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
r=d['roofline']
for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
timeout -k 10 120 python tools/bench_c1.py > $O/r02ad_bench_c1.log 2>&1; cat $O/r02ad_bench_c1.log
