#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02r_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02r_tests.log)"; grep -n "^FAILED" $O/r02r_tests.log | head -30
