#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_pwconv3.py -m gpu -q -x > $O/pw3_tests.log 2>&1; echo "pw3 tests exit=$? $(tail -1 $O/pw3_tests.log)"; grep -n "^FAILED\|Error\|assert " $O/pw3_tests.log | head -20
timeout -k 10 200 python tools/bench_pw.py --iters 10 --modes fwd,res,dgrad,cublas --only layer0 > $O/pw3_bench.log 2>&1
for l in layer1.x layer2.x; do timeout -k 10 100 python tools/bench_pw.py --iters 10 --modes fwd,res,dgrad,cublas --only $l >> $O/pw3_bench.log 2>&1; done; grep layer $O/pw3_bench.log
