#!/usr/bin/env bash
# Round-2 closing evidence (1 GPU) after the BatchNorm apply-pass work: tests, smoke, bench lines of this repo (the reference
# arms are unchanged: profiles/r02zb_bench_*_reference.json), BN op bench, ncu launch list of one eager C3 step.
set -uo pipefail
T=${1:-r02zc}
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/${T}_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/${T}_tests.log)"; grep -n "^FAILED\|^ERROR" $O/${T}_tests.log | head
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "smoke exit=$? $(tail -1 $O/${T}_smoke.log)"
run() { local name=$1; shift; timeout -k 10 400 python bench.py "$@" > $O/${T}_bench_$name.json 2> $O/${T}_bench_$name.err; python - $O/${T}_bench_$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f" % (d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value',0)), d.get('dtype'), d.get('impl'))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run c3_ours --steps 10 --warmup 3
run c4_ours --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline
run c2_ours --tier tiny --dtype fp32 --infer --batch 8 --steps 30 --warmup 5 --no-cpu-baseline
timeout -k 10 200 python tools/bench_bn.py --iters 10 > $O/${T}_bench_bn.log 2>&1; echo "bn bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1300 --csv --log-file $O/${T}_launches_c3.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_launches_c3.log 2>&1; echo launches c3 rc=$?
ls $O | grep ${T}_ | wc -l
