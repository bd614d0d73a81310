#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q -k "attention or aq or AQ or models" > $O/r02ac_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02ac_tests.log)"; grep -n "^FAILED\|Error" $O/r02ac_tests.log | head -20
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ac_bench_c4.json 2> $O/r02ac_bench_c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ac_bench_c4.json').read().strip().splitlines()[-1])
print("C4 value %.1f ms/step %.3f e2e %.1f" % (d['value'], d['ms_per_step'], d['e2e']['value']))
r=d['roofline']
for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1500 --csv --log-file $O/r02ac_launches_c4.csv python bench.py --variant rubiks3d-aq --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/r02ac_launches_c4.log 2>&1; echo launches rc=$?
python tools/launch_summary.py $O/r02ac_launches_c4.csv --by-grid > $O/r02ac_launches_c4_summary.txt; head -45 $O/r02ac_launches_c4_summary.txt
