#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
cap() { local name=$1 rx=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $O/r02z_$name -f "$@" > /dev/null 2>&1; echo "$name rc=$?"; }
cap strip_l3_bwd k_shift3d_strip 2 python tools/prof_case.py --C 288 --H 14 --iters 3
cap strip_l3_fwd k_shift3d_strip 3 python tools/prof_case.py --C 288 --H 14 --iters 3
cap strip_l0_bwd k_shift3d_strip 2 python tools/prof_case.py --C 72 --H 112 --iters 3
cap bn_apply_bwd_l3 "k_bn_apply" 5 python tools/bench_bn.py --only layer3.x --iters 1
ls -la $O/r02z_strip_l3_bwd.ncu-rep $O/r02z_strip_l3_fwd.ncu-rep $O/r02z_strip_l0_bwd.ncu-rep
