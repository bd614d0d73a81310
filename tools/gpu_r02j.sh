#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv2.py -x -q > $O/r02j_pw2_tests.log 2>&1; rc=$?; echo "pw2 tests exit=$rc $(tail -1 $O/r02j_pw2_tests.log)"
for os_ in 2 3 4 6; do for w in 0 20 100 500; do echo "opstages=$os_ waitns=$w"; for d in 31 0; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 --opstages $os_ --waitns $w 2>&1 | grep "^dbg"; done; done; done | tee $O/r02j_sweep_l3.log
for os_ in 2 4; do for w in 0 100; do for kc in 16 32; do echo "opstages=$os_ waitns=$w kc=$kc"; for d in 31 0; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 576 --H 7 --mode fwd --dbg $d --reps 20 --opstages $os_ --waitns $w --kc $kc 2>&1 | grep "^dbg"; done; done; done; done | tee $O/r02j_sweep_l4.log
