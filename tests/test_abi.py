"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/rcbevdet_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "rcbevdet_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rcb_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def handle():
    from rcbevdet_b200 import build
    return ctypes.CDLL(build.build())


def test_header_declares_the_expected_surface():
    syms = _declared_symbols()
    for must in ("rcb_voxel_pooling_prepare_v2", "rcb_bev_pool_v2_fwd", "rcb_bev_pool_v2_bwd",
                 "rcb_radar_rcs_scatter", "rcb_pool_validate", "rcb_planes_to_rows"):
        assert must in syms


def test_library_exports_every_declared_symbol(handle):
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} declared in the header but not exported"


def test_python_binding_covers_every_symbol():
    from rcbevdet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()


def test_struct_layouts_match_header():
    from rcbevdet_b200 import _lib
    assert ctypes.sizeof(_lib.PrepareDesc) == 5 * 4 + 9 * 4
    assert ctypes.sizeof(_lib.PoolDesc) == 15 * 4
    assert ctypes.sizeof(_lib.RadarDesc) == 6 * 4


def test_host_only_entry_points(handle):
    from rcbevdet_b200 import _lib
    lib = _lib.lib()
    assert lib.rcb_version() >= 100
    assert b"workspace" in lib.rcb_error_string(-2)
    d = _lib.PrepareDesc()
    d.B, d.N, d.D, d.H, d.W = 8, 6, 118, 16, 44
    d.lower[:] = [-51.2, -51.2, -5.0]
    d.interval[:] = [0.8, 0.8, 8.0]
    d.size[:] = [128.0, 128.0, 1.0]
    assert lib.rcb_prepare_workspace_bytes(ctypes.byref(d)) >= 2 * 8 * 128 * 128 * 4
    d.size[:] = [128.5, 128.0, 1.0]
    assert lib.rcb_prepare_workspace_bytes(ctypes.byref(d)) == 0     # non-integral grid: unsupported
    # argument errors are reported, not crashed on
    assert lib.rcb_voxel_pooling_prepare_v2(ctypes.byref(d), *([None] * 9), None, 0, 0, None) < 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rcbevdet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_missing_library_fails_loudly(monkeypatch):
    from rcbevdet_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/librcbevdet_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_binding_arity_matches_header():
    """Every ctypes signature has exactly as many parameters as the C declaration."""
    from rcbevdet_b200 import _lib
    with open(os.path.join(ROOT, "include", "rcbevdet_b200.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)\s*;", text)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), f"{name}: header has {n} parameters, binding has {len(args)}"
