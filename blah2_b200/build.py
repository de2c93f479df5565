"""Build the native library in-tree (blah2_b200/lib/libb200dd.so).

nvcc cross-compiles for sm_100a without a GPU; the built .so is git-ignored but
travels with gpurun snapshots.  ``python -m blah2_b200.build`` or ``build_native()``.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200dd.so")

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> str:
    """One object per .cu (compiled in parallel, rebuilt only when it or a shared header changed), then one link."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    shared = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "b200dd.h")]
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and not _newer(obj, [src] + shared):
            return obj, False, ""
        r = subprocess.run([nvcc] + cflags + ["-c", "-o", obj, src], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {os.path.basename(src)}")
        return obj, True, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _, _ in results]
    if verbose:
        sys.stderr.write("".join(log for _, _, log in results))
    if force or any(changed for _, changed, _ in results) or _newer(LIB, objs):
        r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed linking libb200dd.so")
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
