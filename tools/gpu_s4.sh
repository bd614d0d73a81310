#!/usr/bin/env bash
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 600 python -m pytest tests/test_gpu_block.py -m gpu -q -x -k "bn" > $O/s4_bn_tests.log 2>&1; echo "bn tests exit=$? $(tail -1 $O/s4_bn_tests.log)"; grep -n "^FAILED\|^ERROR" $O/s4_bn_tests.log | head
timeout -k 10 200 python tools/bench_bn.py --iters 10 > $O/s4_bench_bn.log 2>&1; python - <<'PY'
import re
for l in open('gpurun_out/s4_bench_bn.log'):
    if not l.startswith('layer'): continue
    name=l.split()[0]
    items=re.findall(r"\| ([a-z+]+(?:\(L2 warm\))?) ([0-9.]+) ms", l)
    print(name, " ".join("%s=%s"%(k,v) for k,v in items if 'warm' not in k))
PY
timeout -k 10 600 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -m gpu -q -x > $O/s4_attn_tests.log 2>&1; echo "attention+block tests exit=$? $(tail -1 $O/s4_attn_tests.log)"; grep -n "^FAILED\|^ERROR" $O/s4_attn_tests.log | head
timeout -k 10 300 python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/s4_bench_c4.json 2> $O/s4_bench_c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench_c4.json').read().strip().splitlines()[-1])
print("C4 value %.1f ms/step %.3f"%(d['value'],d['ms_per_step']))
for k in d['roofline']['all_kernels'][:14]: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/s4_bench_c3.json 2> $O/s4_bench_c3.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/s4_bench_c3.json').read().strip().splitlines()[-1])
print("C3 value %.1f ms/step %.3f e2e %.1f"%(d['value'],d['ms_per_step'],d['e2e']['value']))
for k in d['roofline']['all_kernels'][:12]: print("  %-28s %7.3f ms %5d %7.1f GB/s %.3f" % (k['kernel'][:28],k['kernel_ms_per_step'],k['launches_per_step'],k['achieved'],k['frac']))
PY
