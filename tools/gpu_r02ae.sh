#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > $O/r02ae_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02ae_tests.log)"; grep -n "^FAILED\|Error" $O/r02ae_tests.log | head
timeout -k 10 200 python tools/bench_shift.py --dtype bfloat16 --no-ref > $O/r02ae_bench_shift.log 2>&1; tail -12 $O/r02ae_bench_shift.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ae_bench_c3.json 2> $O/r02ae_bench_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ae_bench_c3.json').read().strip().splitlines()[-1])
print("C3 value %.1f ms/step %.3f e2e %.1f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
r=d['roofline']
for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
