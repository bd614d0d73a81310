#!/usr/bin/env bash
# ncu --set full captures of the 16-bit BatchNorm apply kernel (k_bn_apply_h) in the closing state: backward apply (+ReLU mask)
# on 14x14 and 112x112 maps, forward apply on 112x112.  usage: gpu_cap_bn.sh [tag]
set -uo pipefail
T=${1:-r02zd}
O=gpurun_out; mkdir -p $O
cap() { local name=$1 rx=$2 skip=$3; shift 3
    timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$rx" -s $skip -c 1 -o $O/${T}_$name -f "$@" > $O/${T}_$name.log 2>&1; echo "$name rc=$? $(ls -la $O/${T}_$name.ncu-rep 2>/dev/null | awk '{print $5}')"; }
cap bn_apply_h_bwd_l3 'k_bn_apply_h<.*int.1>' 1 python tools/bench_bn.py --only layer3.x --iters 1
cap bn_apply_h_bwd_l0 'k_bn_apply_h<.*int.1>' 1 python tools/bench_bn.py --only layer0 --iters 1
cap bn_apply_h_fwd_l0 'k_bn_apply_h<.*int.0>' 1 python tools/bench_bn.py --only layer0 --iters 1
