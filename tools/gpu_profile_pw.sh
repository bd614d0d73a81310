#!/bin/bash
# ncu evidence for the tcgen05 pointwise-conv kernels + per-launch list of one bench step (run under gpurun, 1 GPU).
# usage: tools/gpu_profile_pw.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 2000 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1; echo launches rc=$?
for spec in "l0_plain layer0 fwd k_pw_conv" "l3_plain layer3.x fwd k_pw_conv" "l3_bn layer3.x bn k_pw_conv" "l3_dgrad layer3.x dgrad k_pw_conv" \
            "l0_wgrad layer0 wgrad k_pw_wgrad" "l3_wgrad layer3.x wgrad k_pw_wgrad"; do
    set -- $spec
    ncu --set full --clock-control none --import-source on -k regex:$4 -s 3 -c 1 -o gpurun_out/${TAG}_pw_$1 -f \
        python tools/bench_pw.py --only $2 --modes $3 --iters 2 > /dev/null 2>&1; echo "$1 rc=$?"
done
ls -la gpurun_out | grep ${TAG}_
