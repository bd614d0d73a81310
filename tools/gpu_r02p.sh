#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests -m gpu -q -x > $O/r02p_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02p_tests.log)"; grep -n "^FAILED" $O/r02p_tests.log
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
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02p_bench_c3_ours.json 2> $O/r02p_bench_c3_ours.err; summ $O/r02p_bench_c3_ours.json
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02p_bench_c4_ours.json 2> $O/r02p_bench_c4_ours.err; summ $O/r02p_bench_c4_ours.json
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02p_bench_c2_ours.json 2> $O/r02p_bench_c2_ours.err; summ $O/r02p_bench_c2_ours.json
