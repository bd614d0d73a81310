#!/usr/bin/env bash
# Round-2 GPU call B: first run of the second-generation pointwise-conv kernel (parity, op bench), then the rest of the suite.
set -uo pipefail
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02b_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02b_pw2_tests.log)"
if [ $rc -ne 0 ]; then
  timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -q > $O/r02b_pw2_tests_all.log 2>&1; echo "pw2 all exit=$? $(tail -1 $O/r02b_pw2_tests_all.log)"
fi
echo "t=$(( $(date +%s)-T0 ))s"
timeout -k 10 300 python tools/bench_pw.py --modes fwd,fwd2,res,res2,bn,bn2,dgrad,dgrad2,cublas > $O/r02b_bench_pw.log 2>&1; echo "bench_pw exit=$?"; cat $O/r02b_bench_pw.log
echo "t=$(( $(date +%s)-T0 ))s"
timeout -k 10 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_pwconv2.py > $O/r02b_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02b_tests.log)"
echo "t=$(( $(date +%s)-T0 ))s"
if [ $rc -eq 0 ]; then
  timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/r02b_bench_c3_ours.json 2> $O/r02b_bench_c3_ours.err; echo "c3 ours exit=$?"; tail -c 600 $O/r02b_bench_c3_ours.json
fi
echo "t=$(( $(date +%s)-T0 ))s"
