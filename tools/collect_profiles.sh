#!/usr/bin/env bash
# Turns the raw output of an evidence run (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
set -uo pipefail
T=${1:-r02z}; O=gpurun_out; P=profiles
for f in $O/${T}_bench_*.json; do [ -s "$f" ] && tail -1 "$f" > $P/$(basename "$f"); done
for f in bench_pw.log bench_pw_first.log bench_shift.log bench_bn.log sanitizer_summary.txt; do [ -s $O/${T}_$f ] && cp $O/${T}_$f $P/${T}_$f; done
for c in c3 c4; do
  if [ -s $O/${T}_launches_$c.csv ]; then
    { python tools/launch_summary.py $O/${T}_launches_$c.csv --by-grid --title "ncu --metrics gpu__time_duration.sum --clock-control none, one eager $c step (bench.py --graph off), per kernel x grid"
      python tools/launch_summary.py $O/${T}_launches_$c.csv --title "same capture, per kernel"; } > $P/${T}_launches_${c}_summary.txt 2>&1
  fi
done
: > $P/${T}_ncu_kernels.txt
for r in wg3_l0 wg3_l2_bn pw3_l0_bn pw3_l0_res pw_l3_res wg_l3 tiled_s2_bwd; do
  if [ -s $O/${T}_$r.ncu-rep ]; then
    echo "==== ${T}_$r.ncu-rep (ncu --set full --clock-control none --import-source on) ====" >> $P/${T}_ncu_kernels.txt
    python tools/ncu_summary.py $O/${T}_$r.ncu-rep --top 12 >> $P/${T}_ncu_kernels.txt 2>&1
  fi
done
ls -la $P | grep ${T}_ | wc -l
