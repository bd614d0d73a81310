#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_shift3d_tiled -s 2 -c 1 -o $O/r02z_tiled_s2_bwd -f python tools/prof_case.py --C 72 --H 112 --stride 2 --iters 3 > /dev/null 2>&1; echo "cap rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bn_apply -s 3 -c 1 -o $O/r02z_bn_apply_bwd_l3 -f python tools/bench_bn.py --only layer3.x --iters 2 > /dev/null 2>&1; echo "cap rc=$?"
ls -la $O/r02z_tiled_s2_bwd.ncu-rep $O/r02z_bn_apply_bwd_l3.ncu-rep
