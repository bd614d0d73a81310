#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02e_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02e_pw2_tests.log)"
for m in fwd; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode $m; done 2>&1 | grep -v "^   tile [013-5]" > $O/r02e_trace_l3.log; cat $O/r02e_trace_l3.log | cut -c1-400
for m in fwd; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 576 --H 7 --mode $m; done 2>&1 | grep -v "^   tile [013-5]" > $O/r02e_trace_l4.log; cat $O/r02e_trace_l4.log | cut -c1-400
for os_ in 2 3 4; do for kc in 16 32; do echo "opstages=$os_ kc=$kc"; timeout -k 10 120 python tools/bench_pw.py --only layer3.x --modes fwd2,res2,bn2 --opstages $os_ --kc $kc | tail -1; timeout -k 10 120 python tools/bench_pw.py --only layer4.x --modes fwd2,res2,bn2 --opstages $os_ --kc $kc | tail -1; done; done 2>&1 | tee $O/r02e_tuning.log
