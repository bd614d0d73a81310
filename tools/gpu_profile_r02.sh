#!/bin/bash
# Round-2 ncu evidence (run under gpurun, 1 GPU): per-launch list of one eager bench step + --set full captures of the
# dominant kernels.  Numbers printed by anything run under ncu are never bench values.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 2200 --csv --log-file $O/${TAG}_launches_bench.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/${TAG}_launches_bench.log 2>&1; echo launches rc=$?
cap() {  # name kernel-regex command...
    local name=$1 rx=$2; shift 2
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/${TAG}_$name -f "$@" > /dev/null 2>&1; echo "$name rc=$?"
}
cap pw_l3_res k_pw_conv python tools/bench_pw.py --only layer3.x --modes res --iters 2
cap pw_l3_bn k_pw_conv python tools/bench_pw.py --only layer3.x --modes bn --iters 2
cap pw2_l4_res k_pw2 python tools/bench_pw.py --only layer4.x --modes res2 --iters 2
cap pw2_l3_res k_pw2 python tools/bench_pw.py --only layer3.x --modes res2 --iters 2
cap strip_l3_bwd "k_shift3d_strip.*Li1E" python tools/prof_case.py --C 288 --H 14 --iters 2
cap strip_l3_fwd "k_shift3d_strip.*Li0E" python tools/prof_case.py --C 288 --H 14 --iters 2
ls -la $O | grep ${TAG}_
