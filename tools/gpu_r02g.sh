#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02g_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02g_pw2_tests.log)"
for m in fwd res; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode $m; done 2>&1 | grep -v "^   tile [013-5]" > $O/r02g_trace_l3.log; grep -A4 "^CTA 0\|^v2" $O/r02g_trace_l3.log | cut -c1-420
for m in fwd; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 576 --H 7 --mode $m; done 2>&1 | grep -v "^   tile [013-5]" > $O/r02g_trace_l4.log; grep -A4 "^CTA 0\|^v2" $O/r02g_trace_l4.log | cut -c1-420
timeout -k 10 300 python tools/bench_pw.py --modes fwd,fwd2,res,res2,bn,bn2,dgrad2 2>&1 | tee $O/r02g_bench_pw.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02g_bench_c3_ours.json 2> $O/r02g_bench_c3_ours.err; echo "c3 ours exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_c3_ours.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'])
r=d['roofline']
print(r['kernel'][:30], r['kernel_ms_per_step'], r['frac'])
for k in r['all_kernels']: print("%-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
