"""Seeded synthetic scenes for parity tests and benchmarks.

Mirrors gsplat's ``init_type="random"`` initialisation: a cube of random Gaussians
viewed from cameras on a circle, OpenCV convention (+z forward, world->camera
``viewmats``).  Everything is generated on the CPU from ``torch.Generator().manual_seed``
so the CPU oracle and the GPU path see identical bits.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

SH_C0 = 0.2820947917738781


def look_at_viewmat(eye: torch.Tensor, target: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """World->camera 4x4 with +z forward, +x right, +y down."""
    fwd = target - eye
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    R = torch.stack([right, down, fwd], dim=0)          # rows = camera axes in world coords
    V = torch.eye(4, dtype=torch.float32)
    V[:3, :3] = R
    V[:3, 3] = -R @ eye
    return V


def make_cameras(C: int, width: int, height: int, extent: float = 3.0, focal_frac: float = 0.74) -> Dict[str, torch.Tensor]:
    viewmats = []
    for c in range(C):
        ang = 2.0 * math.pi * c / max(C, 1) + 0.3
        eye = torch.tensor([2.3 * extent * math.cos(ang), 2.3 * extent * math.sin(ang), 0.5 * extent])
        viewmats.append(look_at_viewmat(eye, torch.zeros(3), torch.tensor([0.0, 0.0, 1.0])))
    f = focal_frac * width
    K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float32)
    return {"viewmats": torch.stack(viewmats), "Ks": K[None].repeat(C, 1, 1).contiguous()}


def make_scene(N: int, sh_degree: int, C: int, width: int, height: int, seed: int = 0,
               extent: float = 3.0) -> Dict[str, torch.Tensor]:
    """Returns post-activation tensors with the argument names of ``rasterization()``."""
    g = torch.Generator().manual_seed(seed)
    K = (sh_degree + 1) ** 2
    means = extent * (2.0 * torch.rand(N, 3, generator=g) - 1.0)
    base = math.log(0.6 * (8.0 * extent ** 3 / N) ** (1.0 / 3.0))
    scales = torch.exp(base + 0.3 * torch.randn(N, 3, generator=g))
    quats = torch.randn(N, 4, generator=g)
    opacities = torch.sigmoid(1.5 * torch.randn(N, generator=g))
    sh0 = (torch.rand(N, 1, 3, generator=g) - 0.5) / SH_C0
    shN = 0.05 * torch.randn(N, K - 1, 3, generator=g)
    colors = torch.cat([sh0, shN], dim=1).contiguous()
    out = {"means": means, "quats": quats, "scales": scales, "opacities": opacities, "colors": colors}
    out.update(make_cameras(C, width, height, extent))
    return out


def make_scene_blocked(N: int, sh_degree: int, C: int, width: int, height: int, seed: int = 0, extent: float = 3.0,
                       lo: int = 0, hi: int = None, block: int = 1 << 20) -> Dict[str, torch.Tensor]:
    """The same distribution as ``make_scene``, generated in independent blocks of ``block`` Gaussians (block b is
    seeded by (seed, b)), so a rank of a Gaussian-sharded job can build just its shard [lo, hi) of a scene too large
    to generate N times over (30 M Gaussians = 7 GB).  Any [lo, hi) is a slice of the same N-Gaussian scene."""
    hi = N if hi is None else hi
    K = (sh_degree + 1) ** 2
    base = math.log(0.6 * (8.0 * extent ** 3 / N) ** (1.0 / 3.0))
    parts = {k: [] for k in ("means", "quats", "scales", "opacities", "colors")}
    for b in range(lo // block, (max(hi, lo + 1) - 1) // block + 1):
        b0, b1 = b * block, min(N, (b + 1) * block)
        n = b1 - b0
        if n <= 0:
            break
        g = torch.Generator().manual_seed(seed * 1_000_003 + b + 1)
        means = extent * (2.0 * torch.rand(n, 3, generator=g) - 1.0)
        scales = torch.exp(base + 0.3 * torch.randn(n, 3, generator=g))
        quats = torch.randn(n, 4, generator=g)
        opacities = torch.sigmoid(1.5 * torch.randn(n, generator=g))
        sh0 = (torch.rand(n, 1, 3, generator=g) - 0.5) / SH_C0
        shN = 0.05 * torch.randn(n, K - 1, 3, generator=g)
        colors = torch.cat([sh0, shN], dim=1)
        s0, s1 = max(lo, b0) - b0, min(hi, b1) - b0
        for k, v in (("means", means), ("quats", quats), ("scales", scales), ("opacities", opacities), ("colors", colors)):
            parts[k].append(v[s0:s1])
    out = {k: (torch.cat(v).contiguous() if v else torch.empty(0)) for k, v in parts.items()}
    out.update(make_cameras(C, width, height, extent))
    return out
