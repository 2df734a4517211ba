"""Builds libfrg.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m facerecognition_infrenceengine_b200.build [--force] [--verbose]

No torch, no JIT cache: plain ``nvcc`` so that the built ``.so`` sits next to the sources and
travels with the repository snapshot to the GPU box.  cudart is linked statically; the driver API
(cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint, so the library
loads on a machine without libcuda.so (symbol checks) and fails only when compute is requested.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(CSRC, "libfrg.so")
OBJ = os.path.join(CSRC, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
BASE_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function"]
NVCC_FLAGS = BASE_FLAGS + ["-I", INCLUDE, "-I", CSRC]


class NoCompiler(RuntimeError):
    pass


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise NoCompiler("nvcc not found; libfrg.so cannot be built")
    return exe


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(INCLUDE, "frg.h"), "rb").read())
    h.update(" ".join(ARCH + BASE_FLAGS).encode())     # no absolute paths: the tree moves (GPU box)
    return h.hexdigest()


def _fresh(stamp: str, digest: str) -> bool:
    return os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest()
    if not force and _fresh(stamp, digest):
        return LIB
    # one builder at a time (torchrun starts N ranks at once); the others wait, then find it fresh
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _fresh(stamp, digest):
                return LIB
            return _build_locked(stamp, digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, digest: str, verbose: bool) -> str:
    cc = nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ, src[:-3] + ".o")
        cmd = [cc, *ARCH, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    tmp = "%s.tmp.%d" % (LIB, os.getpid())
    cmd = [cc, *ARCH, "-shared", "-o", tmp, *objs, "-Xcompiler", "-fPIC", "-cudart", "static",
           "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)                      # atomic: a concurrent dlopen never sees a partial file
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
