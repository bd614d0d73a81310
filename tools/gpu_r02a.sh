#!/usr/bin/env bash
# Round-2 GPU call A: full GPU test suite (incl. whole-model parity vs baseline/_ref and the drop-in boundary test),
# the C2 / C3 / C4 bench lines for both arms (+ the bf16-autocast reference arm), batch sweep, sanitizer, tiling probe.
set -uo pipefail
O=gpurun_out; mkdir -p $O
T0=$(date +%s)
python -m pytest tests -m gpu -x -q > $O/r02a_tests.log 2>&1; echo "tests exit=$? $(tail -1 $O/r02a_tests.log)"
echo "t=$(( $(date +%s)-T0 ))s"
python __graft_entry__.py --smoke > $O/r02a_smoke.log 2>&1; echo "smoke exit=$? $(tail -1 $O/r02a_smoke.log)"
# tiling probe for the redesign: layer3 / layer4 with 1, 2, 3, 4 output-channel splits
for sp in 0 2 3 4; do echo "splits=$sp"; python tools/bench_pw.py --only layer3.x --modes fwd,res,bn,dgrad --splits $sp; python tools/bench_pw.py --only layer4.x --modes fwd,res,bn,dgrad --splits $sp; done > $O/r02a_pw_splits.log 2>&1
python tools/bench_pw.py > $O/r02a_bench_pw.log 2>&1
echo "t=$(( $(date +%s)-T0 ))s"
# bench lines: ours
python bench.py --steps 10 --warmup 3 > $O/r02a_bench_c3_ours.json 2> $O/r02a_bench_c3_ours.err; echo "c3 ours exit=$?"
python bench.py --variant rubiks3d-aq --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_c4_ours.json 2> $O/r02a_bench_c4_ours.err; echo "c4 ours exit=$?"
python bench.py --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02a_bench_c2_ours.json 2> $O/r02a_bench_c2_ours.err; echo "c2 ours exit=$?"
python bench.py --tier tiny --dtype fp32 --infer --batch 1 --steps 50 --warmup 5 --no-cpu-baseline > $O/r02a_bench_c2n1_ours.json 2> $O/r02a_bench_c2n1_ours.err; echo "c2 n1 ours exit=$?"
for b in 8 16 64; do python bench.py --batch $b --steps 10 --warmup 3 --no-cpu-baseline > $O/r02a_bench_c3_b${b}_ours.json 2> $O/r02a_bench_c3_b${b}_ours.err; echo "c3 b$b exit=$?"; done
echo "t=$(( $(date +%s)-T0 ))s"
# reference arms
python bench.py --impl reference --steps 5 --warmup 3 > $O/r02a_bench_c3_ref.json 2> $O/r02a_bench_c3_ref.err; echo "c3 ref exit=$?"
python bench.py --impl reference --ref-autocast --steps 5 --warmup 3 > $O/r02a_bench_c3_ref_autocast.json 2> $O/r02a_bench_c3_ref_autocast.err; echo "c3 ref autocast exit=$?"
python bench.py --impl reference --variant rubiks3d-aq --steps 5 --warmup 3 > $O/r02a_bench_c4_ref.json 2> $O/r02a_bench_c4_ref.err; echo "c4 ref exit=$?"
python bench.py --impl reference --variant rubiks3d-aq --ref-autocast --steps 5 --warmup 3 > $O/r02a_bench_c4_ref_autocast.json 2> $O/r02a_bench_c4_ref_autocast.err; echo "c4 ref autocast exit=$?"
python bench.py --impl reference --tier tiny --dtype fp32 --infer --batch 8 --steps 20 --warmup 5 > $O/r02a_bench_c2_ref.json 2> $O/r02a_bench_c2_ref.err; echo "c2 ref exit=$?"
echo "t=$(( $(date +%s)-T0 ))s"
timeout 900 bash tools/sanitize.sh r02a
echo "t=$(( $(date +%s)-T0 ))s"
