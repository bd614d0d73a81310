#!/usr/bin/env bash
# Builds the UNMODIFIED reference (StanfordVL/RubiksNet, /root/reference) into baseline/_ref/
# so that the GPU box can run it beside the new kernels (bench.py --impl reference, golden-vector
# generation, GPU parity tests).  Test/bench infrastructure only -- never on the product path.
#
# /root/reference is read-only and setup.py writes a build/ dir, so the install runs from a copy under
# /tmp.  The only source change is the 4-line torch>=2.x compatibility patch named in SURVEY.md
# (cuda_src/rubiks2d_kernels.cu:422,448,477,511  `.type()` -> `.scalar_type()`); without it the
# AT_DISPATCH macro no longer compiles.  Nothing is copied into git history: baseline/_ref is
# git-ignored (but NOT gpurun-ignored, so the built package travels to the GPU box).
set -euo pipefail
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${RUBIKS_REFERENCE:-/root/reference}"
WORK="${TMPDIR:-/tmp}/rubiksnet_ref_build"
DEST="$REPO/baseline/_ref"
if [ ! -d "$SRC" ]; then echo "reference not present at $SRC; keeping prebuilt $DEST" >&2; exit 0; fi
rm -rf "$WORK"; mkdir -p "$WORK"
cp -r "$SRC/cuda_src" "$SRC/rubiksnet" "$SRC/setup.py" "$SRC/README.md" "$WORK/"
sed -i 's/AT_DISPATCH_FLOATING_TYPES_AND_HALF(\([a-z_]*\)\.type()/AT_DISPATCH_FLOATING_TYPES_AND_HALF(\1.scalar_type()/' \
    "$WORK/cuda_src/rubiks2d_kernels.cu"
mkdir -p "$DEST"
cd "$WORK"
export TORCH_CUDA_ARCH_LIST="10.0" MAX_JOBS="${MAX_JOBS:-8}"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --upgrade --target "$DEST" "$WORK" 2>&1 | tail -n 5
# the 9 shipped checkpoints (196 MB) travel with the reference install for whole-model parity tests on the GPU box
mkdir -p "$DEST/pretrained" && cp -n "$SRC"/pretrained/*.pth.tar "$DEST/pretrained/"
ls -la "$DEST"
