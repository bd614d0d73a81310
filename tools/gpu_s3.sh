#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_block.py -m gpu -q -x > $O/s3_parity.log 2>&1; echo "parity exit=$? $(tail -1 $O/s3_parity.log)"; grep -n "^FAILED\|^ERROR" $O/s3_parity.log | head
timeout -k 10 300 python tools/bench_shift.py --iters 10 --no-ref > $O/s3_shift.log 2>&1; grep "s=1" $O/s3_shift.log | cut -c1-100
