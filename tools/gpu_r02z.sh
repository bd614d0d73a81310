#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_pwconv_tf32.py tests/test_gpu_pdl.py -m gpu -q > $O/r02z_tests_tf32.log 2>&1; echo "tf32+pdl tests exit=$? $(tail -1 $O/r02z_tests_tf32.log)"; grep -n "^FAILED" $O/r02z_tests_tf32.log | head -20
for b in 8 1; do
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch $b --steps 30 --warmup 5 --no-cpu-baseline > $O/r02z_bench_c2_b$b.json 2> $O/r02z_bench_c2_b$b.err; python - $b <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r02z_bench_c2_b%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print("C2 batch", sys.argv[1], "value %.1f ms/step %.3f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
except Exception as e: print("C2 unreadable", e)
PY
done
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file $O/r02z_launches_c2.csv \
    python bench.py --tier tiny --dtype fp32 --infer --batch 8 --graph off --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02z_launches_c2.log 2>&1; echo launches rc=$?
python tools/launch_summary.py $O/r02z_launches_c2.csv --by-grid > $O/r02z_launches_c2_summary.txt; head -12 $O/r02z_launches_c2_summary.txt
