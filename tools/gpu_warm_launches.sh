#!/usr/bin/env bash
# ncu launch list of one eager C3 step WITHOUT the cache flush between kernels (--cache-control none): per-kernel times with the
# L2 in the state the previous kernel left it, i.e. what the 14x14 stage (28.9 MB tensors) sees inside a step
set -uo pipefail
O=gpurun_out; mkdir -p $O
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 2300 -c 1300 --csv --log-file $O/warm_launches_c3.csv \
    python bench.py --graph off --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $O/warm_launches_c3.log 2>&1; echo launches rc=$?
python tools/launch_summary.py $O/warm_launches_c3.csv --by-grid --title "ncu --cache-control none, one eager C3 step, per kernel x grid" | head -45
