#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02d_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02d_pw2_tests.log)"
timeout -k 10 300 python tools/bench_pw.py --modes fwd,fwd2,res,res2,bn,bn2,dgrad,dgrad2 > $O/r02d_bench_pw.log 2>&1; echo "bench_pw exit=$?"; cat $O/r02d_bench_pw.log
for m in fwd res bn; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode $m; done > $O/r02d_trace_l3.log 2>&1; cat $O/r02d_trace_l3.log
for m in fwd res; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 576 --H 7 --mode $m; done > $O/r02d_trace_l4.log 2>&1; cat $O/r02d_trace_l4.log
echo "t=$(( $(date +%s)-T0 ))s"
timeout -k 10 900 python -m pytest tests/test_gpu_pwconv.py tests/test_gpu_block.py -m gpu -q > $O/r02d_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02d_tests.log)"; grep -n "^FAILED" $O/r02d_tests.log
echo "t=$(( $(date +%s)-T0 ))s"
if [ $rc -eq 0 ]; then
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02d_bench_c3_ours.json 2> $O/r02d_bench_c3_ours.err; echo "c3 ours exit=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_c3_ours.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
r=d['roofline']
print(r['kernel'][:30], r['kernel_ms_per_step'], r['frac'])
for k in r['all_kernels']: print("%-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
fi
echo "t=$(( $(date +%s)-T0 ))s"
