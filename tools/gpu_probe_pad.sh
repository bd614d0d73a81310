#!/usr/bin/env bash
# what the tensor-map schedules would do on the 14x14 / 7x7 stages at a padded pitch (200 / 56 / 64 elements per channel row)
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tiled or stride" > $O/pad_parity.log 2>&1; echo "parity exit=$? $(tail -1 $O/pad_parity.log)"
: > $O/pad_probe.log
timeout -k 10 120 python tools/bench_pw.py --iters 10 --modes fwd,res,bn,dgrad,wgrad,wgrad_bn,cublas --only layer3.x 2>&1 | grep layer >> $O/pad_probe.log
timeout -k 10 120 python tools/bench_pw.py --iters 10 --modes fwd,res,bn,dgrad,wgrad,wgrad_bn --only pad14 --geom pad14,288,25,8 2>&1 | grep pad14 >> $O/pad_probe.log
timeout -k 10 120 python tools/bench_pw.py --iters 10 --modes wgrad,wgrad_bn --only layer4.x 2>&1 | grep layer >> $O/pad_probe.log
timeout -k 10 120 python tools/bench_pw.py --iters 10 --modes wgrad,wgrad_bn --only pad7 --geom pad7,576,7,8 2>&1 | grep pad7 >> $O/pad_probe.log
timeout -k 10 120 python tools/bench_pw.py --iters 10 --modes wgrad,wgrad_bn --only pad7b --geom pad7b,576,8,8 2>&1 | grep pad7b >> $O/pad_probe.log
cat $O/pad_probe.log
