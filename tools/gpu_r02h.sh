#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02h_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02h_pw2_tests.log)"
for d in 0 1 2 3 4 8 12 15 16 17 19 31; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 2>&1 | grep "^dbg\|kernel span"; done | tee $O/r02h_dbg_l3.log
for d in 0 1 15 16 31; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 576 --H 7 --mode fwd --dbg $d --reps 20 2>&1 | grep "^dbg\|kernel span"; done | tee $O/r02h_dbg_l4.log
timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd 2>&1 | grep -v "^   tile [013-5]" | grep -A4 "^CTA 0\|^v2" | cut -c1-420 | tee $O/r02h_trace_l3.log
timeout -k 10 300 python tools/bench_pw.py --only layer3.x --modes fwd,fwd2,res,res2,bn,bn2,dgrad2 2>&1 | tee $O/r02h_bench_pw.log
