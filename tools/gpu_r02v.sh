#!/usr/bin/env bash
# tf32 kernel: tests + C2 bench; BN op bench; ncu captures of the BN and strip kernels at the layer3 geometry
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_pwconv_tf32.py -m gpu -q -x > $O/r02v_tests_tf32.log 2>&1; echo "tf32 tests exit=$? $(tail -1 $O/r02v_tests_tf32.log)"; grep -n "^FAILED\|Error\|assert" $O/r02v_tests_tf32.log | head -20
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02v_bench_c2.json 2> $O/r02v_bench_c2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_c2.json').read().strip().splitlines()[-1])
    print("C2 value %.1f ms/step %.3f e2e %.1f launches %s" % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    r=d['roofline']; print(" top", r['kernel'][:50], r['kernel_ms_per_step'], r['frac'])
    for k in r['all_kernels']: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
except Exception as e: print("C2 unreadable", e)
PY
tail -5 $O/r02v_bench_c2.err
timeout -k 10 300 python bench.py --tier tiny --dtype fp32 --infer --batch 1 --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > $O/r02v_bench_c2n1.json 2> $O/r02v_bench_c2n1.err; python -c "
import json;d=json.loads(open('gpurun_out/r02v_bench_c2n1.json').read().strip().splitlines()[-1]);print('C2 N=1', d['value'], d['ms_per_step'])"
timeout -k 10 300 python tools/bench_bn.py --iters 10 > $O/r02v_bench_bn.log 2>&1; cat $O/r02v_bench_bn.log
cap() { local name=$1 rx=$2; shift 2
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 1 -o $O/r02v_$name -f "$@" > /dev/null 2>&1; echo "$name rc=$?"; }
cap bn_apply_bwd "k_bn_apply.*Li1E" python tools/bench_bn.py --only layer3.x --iters 2
cap bn_reduce_bwd "k_bn_reduce.*Li1E" python tools/bench_bn.py --only layer3.x --iters 2
cap strip_l3_bwd "k_shift3d_strip" python tools/prof_case.py --C 288 --H 14 --iters 3
ls -la $O | grep r02v_
