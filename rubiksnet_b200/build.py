"""Builds librubiks_b200.so (hand-written CUDA for sm_100a + the C ABI of include/rubiks_b200.h).

The library has no torch / ATen dependency, so it compiles in seconds with plain nvcc and the built
.so lives in-tree (rubiksnet_b200/lib/), which is what travels to the GPU box.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librubiks_b200.so")
SOURCES = ["abi.cu", "shift3d_generic.cu", "shift3d_tiled.cu", "shift3d_strip.cu", "shift2d_generic.cu", "attention_shift.cu", "bn_relu.cu", "se.cu", "pw_conv.cu", "pw_conv2.cu", "pw_conv_tf32.cu", "im2col.cu", "input_norm.cu", "pw_conv3.cu", "pw_wgrad3.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=true",  # no --use_fast_math: IEEE div/sqrt; FMA contraction like the reference build
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--threads", "4",
]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cand = os.path.join(cuda_home, "bin", "nvcc")
    return cand if os.path.exists(cand) else "nvcc"


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(_HERE, "..", "include", "rubiks_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, trace=False):
    """Compile every .cu under csrc/ for sm_100a into one shared library.  Returns its path.
    trace=True builds the DEBUG library librubiks_b200_trace.so (-DRB_DEBUG_TRACE: pipeline time stamps for
    tools/trace_pw.py); the product library never contains that code and nothing in the package loads the debug one."""
    lib_path = LIB_PATH.replace(".so", "_trace.so") if trace else LIB_PATH
    if not trace and not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", "_trace.o" if trace else ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *(["-DRB_DEBUG_TRACE"] if trace else []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building librubiks_b200.so")
    subprocess.check_call([_nvcc(), "-shared", "-o", lib_path, *objs, "-Xcompiler", "-fPIC"])
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
