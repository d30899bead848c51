"""In-tree build of libctgcn_b200.so (hand-written CUDA for sm_100a behind a C-ABI).

    python ctgcn_b200/build.py [--force] [--verbose]     (run by path: the package itself needs the .so)

nvcc cross-compiles without a GPU.  The .so stays next to this file (git-ignored, but it
travels with gpurun snapshots) so that the GPU box loads exactly what was built here.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libctgcn_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("CTGCN_NVCC_FLAGS", "").split()


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths, key=os.path.basename):
        h.update(os.path.basename(p).encode())      # not the absolute path: the stamp must stay valid when the tree is copied
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ctgcn_b200.h")]
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [NVCC, *ARCH, *CFLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log = res.stdout + res.stderr
        with open(os.path.join(OBJ_DIR, src[:-3] + ".log"), "w") as fh:
            fh.write(log)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{log}")
        if verbose:
            print(log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
