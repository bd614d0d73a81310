#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
for d in 31 287 543 799; do timeout -k 10 60 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 2>&1 | grep -v "^   tile [013]" | grep -A7 "^CTA 0\|^v2\|^dbg" | grep -v "stage 3: arrival" | cut -c1-420; done | tee $O/r02m_dbg_l3.log
