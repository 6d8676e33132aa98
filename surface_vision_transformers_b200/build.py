"""Builds the sm_100a shared library (C-ABI) in-tree with nvcc.

The library is linked against the static CUDA runtime only, so it loads on a CPU-only box
(symbol-export tests) and travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsvit_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "--use_fast_math" if False else "-DSVIT_NO_FAST_MATH",
] + (["-DSVIT_SPIN_SLEEP_NS=" + os.environ["SVIT_SPIN_SLEEP_NS"]] if os.environ.get("SVIT_SPIN_SLEEP_NS") else [])


def sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp"))
    )


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in os.listdir(CSRC):
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    inc = os.path.join(os.path.dirname(HERE), "include")
    for f in os.listdir(inc):
        if os.path.getmtime(os.path.join(inc, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    inc = os.path.join(os.path.dirname(HERE), "include")
    cmd = [nvcc] + NVCC_FLAGS + ["-I", inc, "-I", CSRC, "-o", LIB] + sources()
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libsvit_b200.so")
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
