#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 120 python tools/debug_tf32.py > $O/r02w_debug_tf32.log 2>&1; cat $O/r02w_debug_tf32.log | tail -40
timeout -k 10 600 python -m pytest tests/test_gpu_pwconv_tf32.py -m gpu -q > $O/r02w_tests_tf32.log 2>&1; echo "tf32 tests exit=$? $(tail -1 $O/r02w_tests_tf32.log)"; grep -n "^FAILED" $O/r02w_tests_tf32.log | head -20
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02w_bench_c2.json 2> $O/r02w_bench_c2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02w_bench_c2.json').read().strip().splitlines()[-1])
    print("C2 value %.1f ms/step %.3f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    r=d['roofline']; print(" top", r['kernel'][:50], r['kernel_ms_per_step'], r['frac'])
    for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
except Exception as e: print("C2 unreadable", e)
PY
tail -3 $O/r02w_bench_c2.err
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 1 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > $O/r02w_bench_c2n1.json 2> $O/r02w_bench_c2n1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02w_bench_c2n1.json').read().strip().splitlines()[-1]);print('C2 N=1', d['value'], d['ms_per_step'])"
