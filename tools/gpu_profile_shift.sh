#!/bin/bash
# ncu evidence for the shift kernels + the per-launch list of one bench step (run under gpurun, 1 GPU).
# usage: tools/gpu_profile_shift.sh <tag>      outputs: gpurun_out/<tag>_*.{csv,ncu-rep,log}
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 2300 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_launches_bench.log 2>&1; echo launches rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_shift3d_strip -s 4 -c 2 -o gpurun_out/${TAG}_strip_l3_bf16 -f \
    python tools/prof_case.py --C 288 --H 14 --batch 32 > gpurun_out/${TAG}_prof1.log 2>&1; echo p1 rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_shift3d_strip -s 4 -c 2 -o gpurun_out/${TAG}_strip_l0_bf16 -f \
    python tools/prof_case.py --C 72 --H 112 --batch 32 > gpurun_out/${TAG}_prof2.log 2>&1; echo p2 rc=$?
ncu --set full --clock-control none --import-source on -k regex:k_shift3d_tiled -s 4 -c 2 -o gpurun_out/${TAG}_tiled_l1_0_bf16 -f \
    python tools/prof_case.py --C 72 --H 112 --stride 2 --batch 32 > gpurun_out/${TAG}_prof3.log 2>&1; echo p3 rc=$?
ls -la gpurun_out | tail -12
