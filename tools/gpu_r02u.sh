#!/usr/bin/env bash
# programmatic dependent launch: bit-exactness tests, full GPU suite, A/B bench lines (C3 on/off, C4, C2)
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_pdl.py tests/test_gpu_block.py -m gpu -q -x > $O/r02u_tests_pdl.log 2>&1; echo "pdl tests exit=$? $(tail -1 $O/r02u_tests_pdl.log)"; grep -n "^FAILED\|Error" $O/r02u_tests_pdl.log | head -20
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02u_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02u_tests.log)"; grep -n "^FAILED" $O/r02u_tests.log | head -30
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
except Exception as e:
    print(sys.argv[1], "unreadable", e); sys.exit(0)
print(sys.argv[1], "value %.1f ms/step %.2f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0), d.get('gpu_launches')))
PY
}
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --pdl off > $O/r02u_bench_c3_pdl_off.json 2> $O/r02u_bench_c3_pdl_off.err; summ $O/r02u_bench_c3_pdl_off.json
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02u_bench_c3_pdl_on.json 2> $O/r02u_bench_c3_pdl_on.err; summ $O/r02u_bench_c3_pdl_on.json; tail -3 $O/r02u_bench_c3_pdl_on.err
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02u_bench_c4.json 2> $O/r02u_bench_c4.err; summ $O/r02u_bench_c4.json
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02u_bench_c2.json 2> $O/r02u_bench_c2.err; summ $O/r02u_bench_c2.json
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline --pdl off > $O/r02u_bench_c2_pdl_off.json 2> $O/r02u_bench_c2_pdl_off.err; summ $O/r02u_bench_c2_pdl_off.json
