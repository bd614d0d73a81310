#!/usr/bin/env bash
# BatchNorm kernels: GPU tests of the BN / block / model paths, compute-sanitizer memcheck + racecheck over the "bn" group of
# tools/sanitize_cases.py, op bench.  usage: gpu_bn.sh [tag]
set -uo pipefail
T=${1:-r02zc}
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_block.py tests/test_gpu_models_vs_reference.py -m gpu -q -x > $O/${T}_bn_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/${T}_bn_tests.log)"; grep -n "^FAILED\|^ERROR" $O/${T}_bn_tests.log | head
GROUPS_=bn bash tools/sanitize.sh ${T}_bn
timeout -k 10 200 python tools/bench_bn.py --iters 10 > $O/${T}_bench_bn.log 2>&1; echo "bn bench rc=$?"
