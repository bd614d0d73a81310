#!/usr/bin/env bash
# full GPU suite + 2-GPU whole-step graph check (run with gpurun --gpus 2)
set -uo pipefail
O=gpurun_out; mkdir -p $O
timeout -k 10 900 python -m pytest tests -m gpu -q > $O/r02s_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02s_tests.log)"; grep -n "^FAILED" $O/r02s_tests.log | head -30
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
if [ "$NG" -ge 2 ]; then
  timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02s_bench_2gpu_whole.json 2> $O/r02s_bench_2gpu_whole.err; echo "2gpu whole exit=$?"; tail -c 400 $O/r02s_bench_2gpu_whole.err; python -c "
import json;d=json.loads(open('gpurun_out/r02s_bench_2gpu_whole.json').read().strip().splitlines()[-1]);print('whole', d['value'], d['ms_per_step'], d['config']['launch'])"
  timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --graph-multi fwdbwd > $O/r02s_bench_2gpu_fwdbwd.json 2> $O/r02s_bench_2gpu_fwdbwd.err; echo "2gpu fwdbwd exit=$?"; python -c "
import json;d=json.loads(open('gpurun_out/r02s_bench_2gpu_fwdbwd.json').read().strip().splitlines()[-1]);print('fwdbwd', d['value'], d['ms_per_step'], d['config']['launch'])"
fi
