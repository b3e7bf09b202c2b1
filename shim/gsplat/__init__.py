"""Shadow package: put ``<repo>/shim`` in front of ``sys.path`` (or PYTHONPATH) and
``from gsplat.rendering import rasterization`` / ``from gsplat import rasterization`` resolve to the
B200-native implementation, which is how the reference's unchanged Train-Stage1 callers
(nerfstudio ``SplatfactoModel.get_outputs`` behind ``ns-train``, source/container/src/main.py:1308-1316;
``gsplat/examples/simple_trainer.py``, source/container/src/main.py:1339-1347) pick it up.
Only the rasterization path is provided; everything else of gsplat is intentionally absent."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pipeline_pointcloud_b200 import (fully_fused_projection, isect_offset_encode, isect_tiles,  # noqa: E402,F401
                                      quat_scale_to_covar_preci, rasterization, rasterize_to_pixels,
                                      spherical_harmonics)

__version__ = "1.4.0+b200"
