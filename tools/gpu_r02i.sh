#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02i_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02i_pw2_tests.log)"
for d in 0 1 31; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 2>&1 | grep -v "^   tile [013]" | grep -A6 "^CTA 0\|^v2\|^dbg" | cut -c1-420; done | tee $O/r02i_dbg_l3.log
timeout -k 10 300 python tools/bench_pw.py --only layer3.x --modes fwd,fwd2,res,res2,bn,bn2,dgrad2 2>&1 | tee $O/r02i_bench_pw.log
timeout -k 10 300 python tools/bench_pw.py --only layer4.x --modes fwd,fwd2,res,res2,bn,bn2,dgrad2 2>&1 | tee -a $O/r02i_bench_pw.log
