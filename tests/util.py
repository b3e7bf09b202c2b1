"""Helpers shared by the GPU parity tests."""
import torch

from oracle import torch_oracle as O
from pipeline_pointcloud_b200.scene import make_scene

LEAVES = ("means", "quats", "scales", "opacities", "colors")


def scene_args(N, deg, C, W, H, seed=0, device="cpu", requires_grad=False, stress=False):
    s = make_scene(N, deg, C, W, H, seed=seed)
    if stress:  # hit every cull branch: behind the camera, far outside the frustum, huge, tiny
        s["means"][: N // 50] *= 4.0
        s["scales"][N // 50: N // 30] *= 30.0
        s["scales"][N // 30: N // 25] *= 0.01
    out = {}
    for k, v in s.items():
        t = v.to(device)
        if requires_grad and k in LEAVES:
            t = t.clone().requires_grad_()
        out[k] = t
    return out


def call(fn, a, W, H, **kw):
    return fn(a["means"], a["quats"], a["scales"], a["opacities"], a["colors"], a["viewmats"], a["Ks"], W, H, **kw)


def grad_report(got, ref):
    """norm-relative error, and the fraction of elements outside |d| <= 1e-3*|ref| + 1e-3*mean|ref|."""
    got, ref = got.detach().cpu().double(), ref.detach().cpu().double()
    rel = ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()
    tol = 1e-3 * ref.abs() + 1e-3 * ref.abs().mean()
    bad = ((got - ref).abs() > tol).double().mean().item()
    return rel, bad


def report(name, payload):
    """Appends one JSON line to gpurun_out/parity_report.jsonl (the GPU box merges gpurun_out/ back; the summaries that
    are kept live under profiles/)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **payload}) + "\n")
    except OSError:
        pass
