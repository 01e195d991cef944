"""Load pieces of the UNMODIFIED reference from /root/reference by path, with the
few third-party names they need at import time stubbed out.

TEST INFRASTRUCTURE ONLY.  /root/reference exists in the build container, not on
the GPU box, so everything here is used (a) by oracle/gen_golden.py to produce the
committed golden vectors and (b) by CPU tests that skip when the tree is absent.
No reference source is copied: modules are exec'd from where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("RCB_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "mmdet3d/models/necks/view_transformer.py"))


def _pkg(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_pkg(parent), leaf, m)
    return m


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load_view_transformer(bev_pool_v2_op=None):
    """Return the reference's mmdet3d/models/necks/view_transformer.py module.
    `bev_pool_v2_op` is what `from mmdet3d.ops.bev_pool_v2.bev_pool import bev_pool_v2`
    resolves to (the op under test)."""
    import torch

    def _missing(*a, **k):
        raise RuntimeError("bev_pool_v2 stub called")

    _pkg("mmcv.cnn").build_conv_layer = lambda *a, **k: None
    r = _pkg("mmcv.runner")
    r.BaseModule = torch.nn.Module
    r.force_fp32 = lambda *a, **k: (lambda f: f)
    _pkg("mmdet.models.backbones.resnet").BasicBlock = torch.nn.Module
    _pkg("mmdet3d.ops.bev_pool_v2.bev_pool").bev_pool_v2 = bev_pool_v2_op or _missing
    _pkg("mmdet3d.models.builder").NECKS = _Registry()
    _pkg("mmdet3d.models.necks")
    mod = _load("mmdet3d.models.necks.view_transformer", "mmdet3d/models/necks/view_transformer.py")
    if bev_pool_v2_op is not None:
        mod.bev_pool_v2 = bev_pool_v2_op
    return mod


def load_pillar_scatter():
    """Return the reference's mmdet3d/models/middle_encoders/pillar_scatter.py module."""
    g = _load("_rcb_ref_gaussian", "mmdet3d/core/utils/gaussian.py")
    _pkg("mmcv.runner").auto_fp16 = lambda *a, **k: (lambda f: f)
    _pkg("mmdet3d.models.builder").MIDDLE_ENCODERS = _Registry()
    core = _pkg("mmdet3d.core")
    core.draw_heatmap_gaussian = g.draw_heatmap_gaussian
    core.draw_heatmap_gaussian_feat = g.draw_heatmap_gaussian_feat
    _pkg("mmdet3d.models.middle_encoders")
    return _load("mmdet3d.models.middle_encoders.pillar_scatter",
                 "mmdet3d/models/middle_encoders/pillar_scatter.py")


def load_bev_pool_py(ext_module):
    """Return the reference's mmdet3d/ops/bev_pool_v2/bev_pool.py with `from . import
    bev_pool_v2_ext` resolving to `ext_module`."""
    pkg = _pkg("_rcb_ref_bev_pool_pkg")
    pkg.bev_pool_v2_ext = ext_module
    sys.modules["_rcb_ref_bev_pool_pkg.bev_pool_v2_ext"] = ext_module
    return _load("_rcb_ref_bev_pool_pkg.bev_pool", "mmdet3d/ops/bev_pool_v2/bev_pool.py")


def load_temporal_alignment(grid_interval, grid_lower_bound):
    """The reference's BEVDepth4D.gen_grid / shift_feature (mmdet3d/models/detectors/bevdet_rc.py:
    585-657), executed from where they lie: the two method bodies are cut out of the file by line
    (the module itself imports half of mmdet3d) and bound to a stand-in object that carries the
    three attributes they read.  Returns an object with .gen_grid(...) and .shift_feature(...)."""
    import ast
    import textwrap
    import types as _t

    import torch
    import torch.nn.functional as F

    path = os.path.join(REF_ROOT, "mmdet3d/models/detectors/bevdet_rc.py")
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    lines = src.splitlines()
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == "BEVDepth4D_RC" or isinstance(node, ast.ClassDef):
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ("gen_grid", "shift_feature") \
                        and item.name not in found:
                    found[item.name] = textwrap.dedent("\n".join(lines[item.lineno - 1:item.end_lineno]))
    if set(found) != {"gen_grid", "shift_feature"}:
        raise RuntimeError("gen_grid / shift_feature not found in the reference")
    ns = {"torch": torch, "F": F}
    exec(found["gen_grid"], ns)
    exec(found["shift_feature"], ns)
    vt = _t.SimpleNamespace(grid_interval=torch.as_tensor(grid_interval, dtype=torch.float32),
                            grid_lower_bound=torch.as_tensor(grid_lower_bound, dtype=torch.float32))

    class _Stub:
        grid = None
        img_view_transformer = vt
    stub = _Stub()
    stub.gen_grid = _t.MethodType(ns["gen_grid"], stub)
    stub.shift_feature = _t.MethodType(ns["shift_feature"], stub)
    return stub
