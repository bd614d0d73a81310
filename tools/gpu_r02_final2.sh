#!/usr/bin/env bash
# Round-2 final evidence run (1 GPU), state after the tensor-map weight gradient / BN-producer / stride-2 changes: tests, smoke,
# every bench line (ours + reference arms), op-level benches, ncu launch lists and --set full captures, compute-sanitizer.
# Numbers printed by anything run under ncu / the sanitizer are never bench values.
set -uo pipefail
T=${1:-r02z}
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
run c3_ours_tma_off --steps 10 --warmup 3 --no-cpu-baseline --tma off
run c3_reference_autocast --impl reference --ref-autocast --steps 8 --warmup 3
run c3_b8_ours --batch 8 --steps 20 --warmup 5 --no-cpu-baseline
run c3_b16_ours --batch 16 --steps 10 --warmup 3 --no-cpu-baseline
run c3_b64_ours --batch 64 --steps 6 --warmup 3 --no-cpu-baseline
run c4_reference --impl reference --variant rubiks3d-aq --steps 6 --warmup 3
run c4_ours --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline
run c2_reference --impl reference --tier tiny --dtype fp32 --infer --batch 8 --steps 30 --warmup 5
run c2_ours --tier tiny --dtype fp32 --infer --batch 8 --steps 30 --warmup 5 --no-cpu-baseline
run c2n1_ours --tier tiny --dtype fp32 --infer --batch 1 --steps 50 --warmup 10 --no-cpu-baseline
timeout -k 10 300 python tools/bench_pw.py --iters 10 --modes fwd,res,bn,dgrad,wgrad,wgrad_bn,cublas > $O/${T}_bench_pw.log 2>&1; cat $O/${T}_bench_pw.log
timeout -k 10 300 python tools/bench_pw.py --iters 10 --modes fwd,res,bn,dgrad,wgrad,wgrad_bn --no-tma --only layer0 > $O/${T}_bench_pw_first.log 2>&1
for l in layer1.x layer2.x; do timeout -k 10 100 python tools/bench_pw.py --iters 10 --modes fwd,res,bn,dgrad,wgrad,wgrad_bn --no-tma --only $l >> $O/${T}_bench_pw_first.log 2>&1; done; grep layer $O/${T}_bench_pw_first.log
timeout -k 10 300 python tools/bench_shift.py --no-ref > $O/${T}_bench_shift.log 2>&1; tail -22 $O/${T}_bench_shift.log | cut -c1-120
timeout -k 10 200 python tools/bench_bn.py --iters 10 > $O/${T}_bench_bn.log 2>&1; cut -c1-200 $O/${T}_bench_bn.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 1300 --csv --log-file $O/${T}_launches_c3.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_launches_c3.log 2>&1; echo launches c3 rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1400 --csv --log-file $O/${T}_launches_c4.csv \
    python bench.py --variant rubiks3d-aq --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${T}_launches_c4.log 2>&1; echo launches c4 rc=$?
cap() { local name=$1 rx=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/${T}_$name -f "$@" > /dev/null 2>&1; echo "$name rc=$?"; }
cap wg3_l0 "k_wg3" 3 python tools/bench_pw.py --only layer0 --modes wgrad --iters 2
cap wg3_l2_bn "k_wg3" 3 python tools/bench_pw.py --only layer2.x --modes wgrad_bn --iters 2
cap pw3_l0_bn "k_pw3" 3 python tools/bench_pw.py --only layer0 --modes bn --iters 2
cap pw3_l0_res "k_pw3" 3 python tools/bench_pw.py --only layer0 --modes res --iters 2
cap pw_l3_res "k_pw_conv" 3 python tools/bench_pw.py --only layer3.x --modes res --iters 2
cap wg_l3 "k_pw_wgrad" 3 python tools/bench_pw.py --only layer3.x --modes wgrad --iters 2
cap tiled_s2_bwd "k_shift3d_tiled.*Li1E" 2 python tools/prof_case.py --C 72 --H 112 --stride 2 --iters 3
bash tools/sanitize.sh ${T}
ls $O | grep ${T}_ | wc -l
