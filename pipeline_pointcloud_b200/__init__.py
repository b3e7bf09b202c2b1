"""b200-splat: a B200-native (sm_100a) Gaussian-splat rasterizer, drop-in for the ``rasterization()``
hot path of krishan44/pipeline-pointcloud's splat-training stage (which runs it through gsplat).

Public surface mirrors ``gsplat``: ``rasterization`` plus the stage-level operators.
Importing this package does not need a GPU; calling anything does -- there is no CPU path.
"""
from .ops import (fully_fused_projection, isect_offset_encode, isect_tiles, quat_scale_to_covar_preci,
                  rasterize_to_pixels, spherical_harmonics)
from .rendering import rasterization

__all__ = ["rasterization", "fully_fused_projection", "spherical_harmonics", "isect_tiles", "isect_offset_encode",
           "rasterize_to_pixels", "quat_scale_to_covar_preci"]
__version__ = "0.1.0"
