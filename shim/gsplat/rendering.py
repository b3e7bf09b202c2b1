"""``gsplat.rendering`` shadow: only ``rasterization`` (the hot path)."""
from pipeline_pointcloud_b200.rendering import rasterization  # noqa: F401

__all__ = ["rasterization"]
