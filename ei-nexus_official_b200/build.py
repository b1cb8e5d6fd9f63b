"""Build libeinx.so (sm_100a only) in-tree with nvcc.

    python -m ei-nexus_official_b200.build        # not importable by that spelling; use
    python "ei-nexus_official_b200/build.py"      # or __graft_entry__.build()

Objects go to ``build/`` at the repo root, the shared library next to this file so that it
travels with the source snapshot to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libeinx.so")
TORCH_LIB = os.path.join(PKG, "libeinx_torch.so")
TORCH_SRC = os.path.join(CSRC, "torch", "einx_torch.cpp")
OBJ = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libeinx.so can only be built with the CUDA 12.9 toolkit")
    return exe


def _digest(paths):
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu for sm_100a and link libeinx.so; returns its path."""
    srcs = sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(ROOT, "include", "einx.h"))
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest(deps)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


def build_torch_ops(force: bool = False) -> str:
    """Compile csrc/torch/einx_torch.cpp (TORCH_LIBRARY(einx, ...): the registered PyTorch ops over the C ABI) with
    the host compiler against this interpreter's torch headers and link it to libeinx.so (rpath $ORIGIN)."""
    import torch
    from torch.utils import cpp_extension as ce

    build_library()
    stamp = os.path.join(OBJ, "torch_digest.txt")
    h = hashlib.sha256(torch.__version__.encode())
    for p in (TORCH_SRC, os.path.join(ROOT, "include", "einx.h")):
        with open(p, "rb") as f:
            h.update(f.read())
    digest = h.hexdigest()
    if not force and os.path.exists(TORCH_LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return TORCH_LIB
    os.makedirs(OBJ, exist_ok=True)
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:  # older signature
        inc = ce.include_paths(cuda=True)
    libdirs = ce.library_paths()
    # the system compiler torch itself was built against; $CXX in this image is a wrapper that swaps the linker and
    # start files (-B...), and a library linked that way cannot unwind C++ exceptions (measured: TORCH_CHECK segfaults)
    cxx = os.environ.get("EINX_CXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++"))
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *[f"-I{d}" for d in inc], "-I" + os.path.join(ROOT, "include"), TORCH_SRC, "-o", TORCH_LIB,
           *[f"-L{d}" for d in libdirs], "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-L" + PKG, "-leinx",
           "-Wl,-rpath,$ORIGIN", *[f"-Wl,-rpath,{d}" for d in libdirs]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cxx} failed on {TORCH_SRC}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return TORCH_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_ops(force="--force" in sys.argv))
