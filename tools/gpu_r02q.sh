#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests -m gpu -q -x -k "2d or aq or attention or parity" > $O/r02q_tests_2d.log 2>&1; echo "2d tests exit=$? $(tail -1 $O/r02q_tests_2d.log)"; grep -n "^FAILED" $O/r02q_tests_2d.log; tail -30 $O/r02q_tests_2d.log | grep -n "Error\|assert" | head
timeout -k 10 600 python -m pytest tests -m gpu -q > $O/r02q_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02q_tests.log)"; grep -n "^FAILED" $O/r02q_tests.log
summ() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "value %.1f ms/step %.2f e2e %.1f" % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0)))
r=d.get('roofline')
if r:
    print("  top:", r['kernel'][:40], r['kernel_ms_per_step'], r['frac'])
    for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
}
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02q_bench_c4_ours.json 2> $O/r02q_bench_c4_ours.err; summ $O/r02q_bench_c4_ours.json
