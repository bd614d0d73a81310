#!/usr/bin/env bash
# Round-2 closing evidence (1 GPU) after the shift-kernel instruction work: tests, smoke, bench lines of this repo (the reference
# arms are unchanged: profiles/r02z_bench_*_reference*.json), shift op bench, ncu launch list + strip-kernel captures.
set -uo pipefail
T=${1:-r02zb}
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
run c3_reference --impl reference --steps 8 --warmup 3
run c3_ours --steps 10 --warmup 3
run c3_b8_ours --batch 8 --steps 20 --warmup 5 --no-cpu-baseline
run c3_b16_ours --batch 16 --steps 10 --warmup 3 --no-cpu-baseline
run c3_b64_ours --batch 64 --steps 6 --warmup 3 --no-cpu-baseline
run c4_ours --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline
run c2_reference --impl reference --tier tiny --dtype fp32 --infer --batch 8 --steps 30 --warmup 5
run c2_ours --tier tiny --dtype fp32 --infer --batch 8 --steps 30 --warmup 5 --no-cpu-baseline
run c2n1_ours --tier tiny --dtype fp32 --infer --batch 1 --steps 50 --warmup 10 --no-cpu-baseline
timeout -k 10 300 python tools/bench_shift.py --no-ref > $O/${T}_bench_shift.log 2>&1; grep "tiled" $O/${T}_bench_shift.log | cut -c1-100
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1300 --csv --log-file $O/${T}_launches_c3.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_launches_c3.log 2>&1; echo launches c3 rc=$?
cap() { local name=$1 rx=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/${T}_$name -f "$@" > /dev/null 2>&1; echo "$name rc=$?"; }
cap strip_l3_bwd k_shift3d_strip 2 python tools/prof_case.py --C 288 --H 14 --iters 3
cap strip_l0_bwd k_shift3d_strip 2 python tools/prof_case.py --C 72 --H 112 --iters 3
cap strip_l0_fwd k_shift3d_strip 3 python tools/prof_case.py --C 72 --H 112 --iters 3
cap tiled_s2_bwd k_shift3d_tiled 2 python tools/prof_case.py --C 72 --H 112 --stride 2 --iters 3
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 3 python tools/sanitize_cases.py shift > $O/${T}_sanitizer_memcheck_shift.log 2>&1; echo "memcheck shift exit=$? $(grep -E 'ERROR SUMMARY' $O/${T}_sanitizer_memcheck_shift.log | tail -1)"
timeout 200 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 3 python tools/sanitize_cases.py shift > $O/${T}_sanitizer_racecheck_shift.log 2>&1; echo "racecheck shift exit=$? $(grep -E 'RACECHECK SUMMARY' $O/${T}_sanitizer_racecheck_shift.log | tail -1)"
ls $O | grep ${T}_ | wc -l
