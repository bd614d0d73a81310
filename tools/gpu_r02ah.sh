#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02ah_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02ah_tests.log)"; grep -n "^FAILED\|^ERROR" $O/r02ah_tests.log | head -20
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02ah_bench_c4.json 2> $O/r02ah_bench_c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02ah_bench_c4.json').read().strip().splitlines()[-1])
print("C4 value %.1f ms/step %.3f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
PY
tail -3 $O/r02ah_bench_c4.err
