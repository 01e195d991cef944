"""rcbevdet_b200 -- B200-native (sm_100a) camera->BEV view-transformation pooling.

Public surface (mirrors the reference's names):
    bev_pool_v2, QuickCumsumCuda, TRTBEVPoolv2      <- mmdet3d/ops/bev_pool_v2/bev_pool.py
    voxel_pooling_prepare_v2, install               <- LSSViewTransformer.voxel_pooling_prepare_v2
    voxel_pooling_v2                                <- LSSViewTransformer.voxel_pooling_v2 (fused, sync-free)
    voxel_pooling_prepare_from_calib,
    voxel_pooling_v2_from_calib                     <- get_lidar_coor fused in (view_transformer.py:115-157)
    radar_rcs_scatter, PointPillarsScatterRCS       <- mmdet3d/models/middle_encoders/pillar_scatter.py
    depth_context_split, lss_view_transform         <- LSSViewTransformer.forward's split + softmax (view_transformer.py:316-320)
    shift_feature, gen_grid_transform               <- BEVDepth4D.shift_feature / gen_grid (bevdet_rc.py:585-657)
    GraphedViewPool                                 <- voxel_pooling_v2_from_calib replayed from a CUDA graph (inference)
Everything computes in librcbevdet_b200.so (hand-written CUDA behind a C ABI, include/
rcbevdet_b200.h); importing this package does not need a GPU, calling an operator does.
"""
from .bev_pool import QuickCumsumCuda, TRTBEVPoolv2, bev_pool_v2  # noqa: F401
from .prepare import (frustum_axes, install, pack_calib, prepare_async, prepare_from_calib_async,  # noqa: F401
                      voxel_pooling_prepare_from_calib, voxel_pooling_prepare_v2)
from .view_pool import voxel_pooling_v2, voxel_pooling_v2_from_calib  # noqa: F401
from .radar import PointPillarsScatterRCS, radar_rcs_scatter  # noqa: F401
from .temporal import gen_grid_transform, shift_feature  # noqa: F401
from .lift import depth_context_split, lss_view_transform  # noqa: F401
from .graphed import GraphedViewPool  # noqa: F401
from . import strips  # noqa: F401  (strip kernels: strips.set_mode("auto" | "on" | "off" | "chain"))

__version__ = "0.1.0"
