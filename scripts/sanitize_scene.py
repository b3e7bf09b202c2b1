"""Small end-to-end exercise of every single-GPU kernel of the path, for `compute-sanitizer --tool memcheck|racecheck|initcheck`
(race detection / sanitizers).  Run:  compute-sanitizer --tool memcheck python scripts/sanitize_scene.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pipeline_pointcloud_b200 as b2s  # noqa: E402
from pipeline_pointcloud_b200.scene import make_scene  # noqa: E402
from pipeline_pointcloud_b200.training import FusedAdam, l1_ssim_loss  # noqa: E402

LEAVES = ("means", "quats", "scales", "opacities", "colors")


def run(N, deg, C, W, H, packed, mode, render_mode="RGB", absgrad=False, view_grad=False, binning="coarse"):
    os.environ["B2S_BINNING"] = binning
    s = make_scene(N, deg, C, W, H, seed=3)
    g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in s.items()}
    if view_grad:
        g["viewmats"].requires_grad_()
    r, a, m = b2s.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], g["viewmats"], g["Ks"], W, H,
                                sh_degree=deg, packed=packed, rasterize_mode=mode, render_mode=render_mode, absgrad=absgrad)
    m["means2d"].retain_grad()
    target = torch.rand_like(r[..., :3])
    loss = l1_ssim_loss(r[..., :3].contiguous(), target) + a.mean()
    loss.backward()
    assert m["means2d"].grad is not None
    _ = m["isect_ids"], m["flatten_ids"], m["isect_offsets"], m["radii"]
    opt = FusedAdam([g[k] for k in LEAVES], lr=1e-3)
    opt.step()
    torch.cuda.synchronize()
    return float(loss)


if __name__ == "__main__":
    print(run(3000, 3, 2, 96, 64, False, "antialiased"))                      # TMA coefficient path, dense, two-level binning
    print(run(3000, 3, 1, 96, 64, True, "classic", absgrad=True))             # packed meta, absgrad
    print(run(2000, 1, 2, 80, 48, False, "antialiased", render_mode="RGB+ED", view_grad=True))   # LDG tile path, camera gradients
    print(run(2000, 3, 1, 64, 48, False, "classic", binning="tiles"))         # reference-shaped lists
    print("sanitize scene ok")
