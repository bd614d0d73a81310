#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 150 --csv --log-file $O/r02x_launches_c2.csv \
    python bench.py --tier tiny --dtype fp32 --infer --batch 8 --graph off --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > $O/r02x_launches_c2.log 2>&1; echo launches rc=$?
python tools/launch_summary.py $O/r02x_launches_c2.csv --by-grid | head -40
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pw_tf32 -s 2 -c 1 -o $O/r02x_tf32_l0 -f python bench.py --tier tiny --dtype fp32 --infer --batch 8 --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1; echo "cap rc=$?"
