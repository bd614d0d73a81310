#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
for d in 31 63 127 32 96 0; do timeout -k 10 120 python tools/trace_pw.py --v2 --C 288 --H 14 --mode fwd --dbg $d --reps 20 2>&1 | grep -v "^   tile [013]" | grep -A5 "^CTA 0\|^v2\|^dbg" | grep -v "raw loads" | cut -c1-420; done | tee $O/r02k_dbg_l3.log
