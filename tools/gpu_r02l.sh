#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02l_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02l_pw2_tests.log)"; if [ $rc -ne 0 ]; then tail -30 $O/r02l_pw2_tests.log; fi
for d in 31 0; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 2>&1 | grep -v "^   tile [013]" | grep -A7 "^CTA 0\|^v2\|^dbg" | cut -c1-420; done | tee $O/r02l_dbg_l3.log
for os_ in 2 3 4; do echo "opstages=$os_"; timeout -k 10 120 python tools/bench_pw.py --only layer3.x --modes fwd2,res2,bn2 --opstages $os_ | tail -1;  timeout -k 10 120 python tools/bench_pw.py --only layer4.x --modes fwd2,res2,bn2 --opstages $os_ | tail -1; done 2>&1 | tee $O/r02l_tuning.log
timeout -k 10 300 python tools/bench_pw.py --modes fwd,fwd2,res,res2,bn,bn2,dgrad2 2>&1 | tee $O/r02l_bench_pw.log
