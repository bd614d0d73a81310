#!/usr/bin/env bash
# compute-sanitizer memcheck + racecheck (+ synccheck) over tools/sanitize_cases.py; summaries -> gpurun_out/<tag>_sanitizer_*.log
set -uo pipefail
TAG="${1:-r02}"
OUT=gpurun_out
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  for grp in ${GROUPS_:-pw shift misc bn}; do
    log="$OUT/${TAG}_sanitizer_${tool}_${grp}.log"
    timeout 200 "$CS" --tool "$tool" --print-limit 20 --error-exitcode 3 python tools/sanitize_cases.py "$grp" > "$log" 2>&1
    echo "$tool $grp exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1)"
  done
done 2>&1 | tee "$OUT/${TAG}_sanitizer_summary.txt"
