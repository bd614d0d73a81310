#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 200 python -m pytest tests/test_gpu_pwconv2.py -x -q 2>&1 | tail -1
for kc in 32 64; do for os_ in 2 3; do echo "kc=$kc opstages=$os_"; timeout -k 10 60 python tools/bench_pw.py --only layer3.x --modes fwd2,res2,bn2 --opstages $os_ --kc $kc | tail -1;  timeout -k 10 60 python tools/bench_pw.py --only layer4.x --modes fwd2,res2,bn2 --opstages $os_ --kc $kc | tail -1; done; done 2>&1 | tee $O/r02o_tuning.log
