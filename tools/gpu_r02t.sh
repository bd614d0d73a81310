#!/usr/bin/env bash
# round-2 evidence call: full GPU suite, op-level pw bench with the cuBLAS column, ncu launch list + --set full captures
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02t_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02t_tests.log)"; grep -n "^FAILED" $O/r02t_tests.log | head -30
timeout -k 10 300 python tools/bench_pw.py --iters 10 --modes fwd,fwd2,res,res2,bn,bn2,dgrad,dgrad2,wgrad,cublas > $O/r02t_bench_pw.log 2>&1; cat $O/r02t_bench_pw.log
timeout -k 10 200 python tools/bench_shift.py > $O/r02t_bench_shift.log 2>&1; tail -20 $O/r02t_bench_shift.log
timeout -k 10 900 bash tools/gpu_profile_r02.sh r02t
