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


def _source_hash():
    """Content hash of everything the library is built from (file times do not survive a repo snapshot reliably)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    inc = os.path.join(os.path.dirname(HERE), "include")
    for d in (CSRC, inc):
        for f in sorted(os.listdir(d)):
            if f.endswith((".cu", ".cpp", ".cuh", ".h")):
                with open(os.path.join(d, f), "rb") as fh:
                    h.update(f.encode() + b"\0" + fh.read())
    return h.hexdigest()


def needs_build():
    """True if the library is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".srchash"):
        return True
    with open(LIB + ".srchash") as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compiles the library if it is missing or older than csrc/ / include/.  Safe under torchrun: the build runs under
    an exclusive file lock (the other ranks wait, re-check and find it fresh) and nvcc writes to a temporary file that
    is renamed into place, so no process can dlopen a half-written library."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another rank built it while we waited for the lock
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            if not os.path.exists(nvcc):
                raise RuntimeError(f"libsvit_b200.so is missing or stale and {nvcc} does not exist: cannot build the "
                                   "sm_100a library (there is no fallback path)")
            inc = os.path.join(os.path.dirname(HERE), "include")
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + ["-I", inc, "-I", CSRC, "-o", tmp] + sources()
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libsvit_b200.so")
            os.replace(tmp, LIB)
            with open(LIB + ".srchash", "w") as fh:
                fh.write(_source_hash())
            if verbose:
                print(r.stdout + r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
