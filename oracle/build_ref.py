"""Compile the reference's OWN bev_pool_v2 CUDA extension, unmodified, from the sources where
they lie under /root/reference (mmdet3d/ops/bev_pool_v2/src/bev_pool.cpp, bev_pool_cuda.cu), into
oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  TEST INFRASTRUCTURE ONLY.

It is used (a) by tests/test_gpu_vs_reference_cuda.py to compare this repo's kernels with the
reference kernels on the GPU, and (b) optionally by bench.py to time the reference kernels on
the same inputs.  No reference source is copied: nvcc reads the files in place.  On the GPU box
/root/reference does not exist; `load()` then only loads the prebuilt shared object.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RCB_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "bev_pool_v2_ext"
SRC = [os.path.join(REF_ROOT, "mmdet3d/ops/bev_pool_v2/src", f) for f in ("bev_pool.cpp", "bev_pool_cuda.cu")]


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def sources_available():
    return all(os.path.isfile(s) for s in SRC)


def build_if_possible(verbose=False):
    """Build oracle/_ref/bev_pool_v2_ext.so when the reference tree is present; no-op otherwise."""
    if os.path.isfile(so_path()):
        return so_path()
    if not sources_available():
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils import cpp_extension
    cpp_extension.load(name=NAME, sources=SRC, build_directory=OUT_DIR, verbose=verbose,
                       extra_cuda_cflags=["-O3", "-lineinfo"], is_python_module=False)
    return so_path() if os.path.isfile(so_path()) else None


def load():
    """The reference's pybind module (bev_pool_v2_forward / bev_pool_v2_backward), or None."""
    path = so_path()
    if not os.path.isfile(path):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    if NAME in sys.modules:
        return sys.modules[NAME]
    loader = importlib.machinery.ExtensionFileLoader(NAME, path)
    spec = importlib.util.spec_from_loader(NAME, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[NAME] = mod
    return mod


if __name__ == "__main__":
    print(build_if_possible(verbose=True))
