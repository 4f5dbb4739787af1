"""Builds archi_b200/lib/libarchi_b200.so in-tree with nvcc for sm_100a (no other target), and the
small host-only helper archi_b200/lib/libarchi_text.so (gcc; tokenising for the lexical index).

``python -m archi_b200.build [--force] [--verbose]``; also called by ``__graft_entry__.build()``.
The built .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libarchi_b200.so")
HOSTSRC = os.path.join(HERE, "hostsrc")
TEXT_LIB = os.path.join(LIBDIR, "libarchi_text.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libarchi_b200.so cannot be built")
    return nvcc


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "archi_b200.h")]
    return max(os.path.getmtime(f) for f in files)


def is_stale() -> bool:
    return not os.path.exists(LIB) or os.path.getmtime(LIB) < _deps_mtime()


def build_host(force: bool = False) -> str:
    """libarchi_text.so: plain C, no CUDA (archi_b200/hostsrc/text_index.c)."""
    src = os.path.join(HOSTSRC, "text_index.c")
    if not force and os.path.exists(TEXT_LIB) and os.path.getmtime(TEXT_LIB) >= os.path.getmtime(src):
        return TEXT_LIB
    os.makedirs(LIBDIR, exist_ok=True)
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        raise RuntimeError("gcc not found; libarchi_text.so cannot be built")
    res = subprocess.run([gcc, "-O2", "-shared", "-fPIC", "-o", TEXT_LIB, src], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("building libarchi_text.so failed")
    return TEXT_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    build_host(force)
    if not force and not is_stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= _deps_mtime():
            return obj
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                 "-cudart", "static", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking libarchi_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
