#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/s2_parity.log 2>&1; echo "parity exit=$? $(tail -1 $O/s2_parity.log)"
timeout -k 10 300 python tools/bench_shift.py --dtype bfloat16 --iters 10 --no-ref > $O/s2_shift.log 2>&1; grep "s=2" $O/s2_shift.log | cut -c1-110
timeout -k 10 300 python tools/bench_shift.py --dtype float32 --iters 10 --no-ref > $O/s2_shift32.log 2>&1; grep "s=2" $O/s2_shift32.log | cut -c1-110
