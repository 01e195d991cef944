"""Builds librcbevdet_b200.so (the C-ABI CUDA library, sm_100a only) in-tree with nvcc.

    python -m rcbevdet_b200.build [--force] [--verbose]

The library is plain CUDA C++ with `extern "C"` entry points (include/rcbevdet_b200.h); it does
not link against torch.  nvcc cross-compiles without a GPU, so this runs in the build container;
the resulting .so travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "librcbevdet_b200.so")
OBJ_DIR = os.path.join(PKG, "build")

SOURCES = ["api.cu", "prepare.cu", "prepare_lsd.cu", "pool_plan.cu", "pool_fwd.cu", "pool_fwd_cells.cu", "pool_bwd.cu", "strips.cu", "layout.cu", "radar.cu", "bev_shift.cu", "depth_context.cu", "trt_plugin.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "--expt-extended-lambda", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
    # bit-exact integer parity needs IEEE fp32 sub/div and no flush-to-zero
    "--fmad=true", "--prec-div=true", "--ftz=false",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: cannot build librcbevdet_b200.so")
    return cand


def _deps():
    deps = [os.path.join(ROOT, "include", "rcbevdet_b200.h"), os.path.join(CSRC, "common.cuh"),
            os.path.join(CSRC, "prepare_common.cuh")]
    return deps + [os.path.join(CSRC, s) for s in SOURCES]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    log = r.stdout + r.stderr
    with open(os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".ptxas.log"), "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, verbose), SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
