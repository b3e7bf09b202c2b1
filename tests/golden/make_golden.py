"""Generates tests/golden/scene_small.npz from the oracle (run: python -m tests.golden.make_golden).

The reference holds no golden vector for the rasterization path, and gsplat cannot be
imported here, so the committed vectors are outputs of oracle/torch_oracle.py on a seeded synthetic scene.
They pin (a) the oracle against accidental change and (b) the CUDA path on the GPU box.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import torch_oracle as O  # noqa: E402
from pipeline_pointcloud_b200.scene import make_scene  # noqa: E402

SPEC = dict(N=600, sh_degree=3, C=2, width=72, height=56, seed=11, rasterize_mode="antialiased")


def compute(spec):
    s = make_scene(spec["N"], spec["sh_degree"], spec["C"], spec["width"], spec["height"], seed=spec["seed"])
    leaves = {k: s[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")}
    render, alpha, meta = O.rasterization(
        leaves["means"], leaves["quats"], leaves["scales"], leaves["opacities"], leaves["colors"], s["viewmats"], s["Ks"],
        spec["width"], spec["height"], sh_degree=spec["sh_degree"], packed=False, rasterize_mode=spec["rasterize_mode"],
        return_fragile=True)
    g = torch.Generator().manual_seed(spec["seed"] + 1)
    v_render = torch.rand(render.shape, generator=g) * (~meta["fragile"])[..., None]
    v_alpha = torch.rand(alpha.shape, generator=g) * (~meta["fragile"])[..., None]
    ((render * v_render).sum() + (alpha * v_alpha).sum()).backward()
    out = {
        "render": render.detach().numpy(), "alpha": alpha.detach().numpy(), "fragile": meta["fragile"].numpy(),
        "radii": meta["radii"].numpy(), "means2d": meta["means2d"].detach().numpy(),
        "depths": meta["depths"].detach().numpy(), "conics": meta["conics"].detach().numpy(),
        "tiles_per_gauss": meta["tiles_per_gauss"].numpy(), "isect_ids": meta["isect_ids"].numpy(),
        "flatten_ids": meta["flatten_ids"].numpy(), "isect_offsets": meta["isect_offsets"].numpy(),
        "v_render": v_render.numpy(), "v_alpha": v_alpha.numpy(),
    }
    for k, t in leaves.items():
        out["grad_" + k] = t.grad.numpy()
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scene_small.npz")
    np.savez_compressed(path, **compute(SPEC))
    print("wrote", path, os.path.getsize(path), "bytes")
