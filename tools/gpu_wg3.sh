#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_wgrad3.py -m gpu -q -x > $O/wg3_tests.log 2>&1; echo "wg3 tests exit=$? $(tail -1 $O/wg3_tests.log)"; grep -n "^FAILED\|Error\|assert " $O/wg3_tests.log | head -20
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > $O/wg3_parity.log 2>&1; echo "parity exit=$? $(tail -1 $O/wg3_parity.log)"; grep -n "^FAILED\|Error\|assert " $O/wg3_parity.log | head -20
rm -f $O/wg3_bench.log
for wg in 1,0,8 2,0,8 4,0,8 1,1,8 2,1,8 1,0,4; do
  echo "knobs burst,l2_256,max_stages = $wg" >> $O/wg3_bench.log
  for l in layer0 layer1.x layer2.x; do timeout -k 10 100 python tools/bench_pw.py --iters 10 --modes wgrad,wgrad_bn --only $l --wg $wg 2>&1 | grep layer >> $O/wg3_bench.log; done
done
cat $O/wg3_bench.log
timeout -k 10 300 python tools/bench_shift.py --dtype bfloat16 --iters 10 --no-ref > $O/wg3_shift.log 2>&1; grep "s=2" $O/wg3_shift.log | cut -c1-110
