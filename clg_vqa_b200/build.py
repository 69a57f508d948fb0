"""Build libclgvqa.so (all CUDA kernels + the C ABI) in-tree for sm_100a with nvcc.

    python -m clg_vqa_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting clg_vqa_b200/lib/libclgvqa.so travels to the GPU
box with the repo snapshot (it is git-ignored, not gpurun-ignored).  Objects are compiled in
parallel and cached by a hash of the sources they include.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libclgvqa.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA library cannot be built")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_hash() -> str:
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cuh", ".h")):
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "clgvqa.h"), "rb").read())
    h.update(" ".join(ARCH_FLAGS + NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc, src, hdr_hash, force, verbose):
    path = os.path.join(CSRC, src)
    digest = hashlib.sha256(open(path, "rb").read() + hdr_hash.encode()).hexdigest()[:16]
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    stamp = obj + ".hash"
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == digest:
        return obj, False
    flags = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    cmd = [nvcc] + ARCH_FLAGS + flags + ["-I", os.path.join(ROOT, "include"), "-c", path, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    with open(stamp, "w") as f:
        f.write(digest)
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    hdr_hash = _headers_hash()
    srcs = _sources()
    objs, rebuilt = [], False
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, did in ex.map(lambda s: _compile_one(nvcc, s, hdr_hash, force, verbose), srcs):
            objs.append(obj)
            rebuilt = rebuilt or did
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc] + ARCH_FLAGS + ["-shared", "-Xcompiler", "-fPIC", "-o", LIB] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose=True)
    print("built", lib)
