"""Build liblicv_b200.so in-tree with nvcc for sm_100a (no torch dependency).

    python -m licv_vqa_b200.build [--force] [--verbose]

The shared library is a plain C-ABI library (include/licv_b200.h); nvcc cross-compiles it on a box
without a GPU.  It is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "liblicv_b200.so")
STAMP = os.path.join(LIB_DIR, "liblicv_b200.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-DLICV_BUILD=1",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liblicv_b200 can only be built with the CUDA toolkit")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    files = _sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)
                                if f.endswith((".cuh", ".h")))
    files += [os.path.join(INCLUDE, "licv_b200.h"), os.path.abspath(__file__)]
    h.update(os.environ.get("LICV_EXTRA_NVCC_FLAGS", "").encode())
    for f in files:
        h.update(f.encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into lib/liblicv_b200.so (skipped when sources unchanged)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    for src in _sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        extra = os.environ.get("LICV_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DLICV_TRACE (debug)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-I", CSRC, "-c", src, "-o", obj]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                 stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, cmd, p in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}: {' '.join(cmd)}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        sys.stderr.write("\n".join(log))
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


# -------------------------------------------------------------------------------------------------
# liblicv_torch.so: the TORCH_LIBRARY registration (csrc/licv_torch_ops.cpp), a host-only shim that
# links liblicv_b200.so and libtorch.  g++ only (no kernel in it); one translation unit.
# -------------------------------------------------------------------------------------------------
TORCH_SRC = os.path.join(CSRC, "licv_torch_ops.cpp")
TORCH_LIB = os.path.join(LIB_DIR, "liblicv_torch.so")
TORCH_STAMP = os.path.join(LIB_DIR, "liblicv_torch.stamp")


def _torch_digest():
    import torch
    h = hashlib.sha256()
    h.update(torch.__version__.encode())
    for f in (TORCH_SRC, os.path.join(INCLUDE, "licv_b200.h"), os.path.abspath(__file__)):
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build_torch_ops(force: bool = False) -> str:
    """Compile csrc/licv_torch_ops.cpp into lib/liblicv_torch.so against the installed torch."""
    build()    # the C-ABI library it links
    digest = _torch_digest()
    if not force and os.path.exists(TORCH_LIB) and os.path.exists(TORCH_STAMP):
        with open(TORCH_STAMP) as f:
            if f.read().strip() == digest:
                return TORCH_LIB
    import torch
    from torch.utils import cpp_extension as ce
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Wno-unused-function",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *[f"-I{p}" for p in ce.include_paths()], f"-I{cuda_home}/include", f"-I{INCLUDE}",
           TORCH_SRC, "-o", TORCH_LIB,
           f"-L{LIB_DIR}", "-llicv_b200", f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError(f"g++ failed on {TORCH_SRC}: {' '.join(cmd)}")
    with open(TORCH_STAMP, "w") as f:
        f.write(digest)
    return TORCH_LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    print(build_torch_ops(force="--force" in sys.argv))
