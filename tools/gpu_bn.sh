#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_block.py tests/test_gpu_models_vs_reference.py -m gpu -q -x > $O/bn_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/bn_tests.log)"; grep -n "^FAILED\|^ERROR" $O/bn_tests.log | head
timeout -k 10 200 python tools/bench_bn.py --iters 10 > $O/bn_bench.log 2>&1; python - <<'PY'
import re
for l in open('gpurun_out/bn_bench.log'):
    if not l.startswith('layer'): continue
    name=l.split()[0]
    items=re.findall(r"\| ([a-z+]+(?:\(L2 warm\))?) ([0-9.]+) ms", l)
    print(name, " ".join("%s=%s"%(k,v) for k,v in items if 'warm' not in k))
PY
