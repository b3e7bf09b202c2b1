"""Sub-frustum checker for full-size scenes  --  TEST INFRASTRUCTURE ONLY (same rules as torch_oracle.py: used by
``tests/`` and by the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, never by the product).

A BASELINE-size scene (5.8 M Gaussians, 1297x840) is far too large for the CPU oracle as a whole, but a window of tiles
is not: the oracle projects and bins, on the WHOLE frame, the Gaussians that can reach the window (a conservative
superset, chosen by a cheap bound), composites only the window's tiles, and differentiates a cotangent that is zero
outside the window.  Because the frame geometry is untouched, the window's slice of the sorted (camera, tile, depth)
key list is bit-comparable with the slice of the CUDA path's full-frame list, the window's pixels are comparable at
the 1e-4 bar and the gradients of the selected Gaussians at the 1e-3 bar.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from . import torch_oracle as O

LEAVES = ("means", "quats", "scales", "opacities", "colors")


def central_window(W: int, H: int, win_w: int = 160, win_h: int = 96, tile: int = 16) -> Tuple[int, int, int, int]:
    """Tile rectangle (tx0, ty0, tx1, ty1) of a tile-aligned window of about win_w x win_h pixels in the image centre."""
    tw, th = math.ceil(W / tile), math.ceil(H / tile)
    nx, ny = min(tw, max(1, win_w // tile)), min(th, max(1, win_h // tile))
    tx0, ty0 = (tw - nx) // 2, (th - ny) // 2
    return tx0, ty0, tx0 + nx, ty0 + ny


def select_window_inputs(scene: Dict[str, torch.Tensor], W: int, H: int, window, cam: int = 0, tile: int = 16):
    """Gaussians that can touch the window's tiles, by a conservative bound on the projected radius:
    |J| <= (f/z) sqrt(1 + lim_x^2 + lim_y^2) with the clamped x/z, y/z of the projection (lim = 1.3 tan(fov/2)), so
    radius = ceil(3 sqrt(lambda_max)) <= 3 |J| s_max + 3 sqrt(0.3) + 1.  Returns (sub-scene with camera `cam` only,
    index of the kept Gaussians)."""
    V, K = scene["viewmats"][cam], scene["Ks"][cam]
    tx0, ty0, tx1, ty1 = window
    x0, y0, x1, y1 = tx0 * tile, ty0 * tile, min(W, tx1 * tile), min(H, ty1 * tile)
    t = scene["means"] @ V[:3, :3].T + V[:3, 3]
    z = t[:, 2].clamp_min(1e-6)
    fx, fy = float(K[0, 0]), float(K[1, 1])
    limx = 1.3 * max(W - float(K[0, 2]), float(K[0, 2])) / fx
    limy = 1.3 * max(H - float(K[1, 2]), float(K[1, 2])) / fy
    jn = math.sqrt(1.0 + limx * limx + limy * limy) * max(fx, fy)
    px, py = fx * t[:, 0] / z + K[0, 2], fy * t[:, 1] / z + K[1, 2]
    rad = 3.0 * jn * scene["scales"].max(dim=1).values / z + 4.0
    keep = (t[:, 2] > 0.01) & (px + rad > x0) & (px - rad < x1) & (py + rad > y0) & (py - rad < y1)
    idx = keep.nonzero(as_tuple=True)[0]
    sub = {k: scene[k][idx].clone() for k in LEAVES}
    sub["viewmats"], sub["Ks"] = V[None].clone(), K[None].clone()
    return sub, idx


def window_cotangents(W: int, H: int, window, seed: int = 1, tile: int = 16):
    """Seeded random cotangents of the full image, zero outside the window."""
    tx0, ty0, tx1, ty1 = window
    g = torch.Generator().manual_seed(seed)
    v_r, v_a = torch.rand(1, H, W, 3, generator=g), torch.rand(1, H, W, 1, generator=g)
    m = torch.zeros(1, H, W, 1)
    m[:, ty0 * tile:ty1 * tile, tx0 * tile:tx1 * tile] = 1.0
    return v_r * m, v_a * m


def oracle_window_step(sub, W: int, H: int, sh_degree, mode: str, window, v_r, v_a, dtype=torch.float32,
                       keep_mask=None, zero_fragile: bool = False):
    """One forward + backward of the oracle on the window.  Returns a dict with the window's render/alpha, the fragile
    mask, the filtered key lists, the gradients of the sub-scene's leaves and the cotangents actually used.
    ``keep_mask`` [1,H,W,1] multiplies the cotangents; ``zero_fragile`` zeroes them on the knife-edge pixels this
    forward reports, so that both sides of a parity check differentiate the same function."""
    leaves = [sub[k].detach().clone().to(dtype).requires_grad_() for k in LEAVES]
    r, a, m = O.rasterization(*leaves, sub["viewmats"].to(dtype), sub["Ks"].to(dtype), W, H, sh_degree=sh_degree, packed=False,
                              rasterize_mode=mode, return_fragile=True, tile_window=window)
    vr, va = v_r.to(dtype), v_a.to(dtype)
    if keep_mask is not None:
        vr, va = vr * keep_mask.to(dtype), va * keep_mask.to(dtype)
    if zero_fragile:
        keep = (~m["fragile"])[..., None].to(dtype)
        vr, va = vr * keep, va * keep
    torch.autograd.backward([r, a], [vr, va])
    return {"render": r.detach(), "alpha": a.detach(), "fragile": m["fragile"], "isect_ids": m["isect_ids"],
            "flatten_ids": m["flatten_ids"], "grads": {k: t.grad for k, t in zip(LEAVES, leaves)},
            "n_isects": int(m["isect_ids"].numel()), "v_render": vr.float(), "v_alpha": va.float()}


def window_key_slice(isect_ids: torch.Tensor, flatten_ids: torch.Tensor, tw: int, th: int, window):
    """The slice of a full-frame sorted list (single camera) that belongs to the window's tiles."""
    tx0, ty0, tx1, ty1 = window
    tile_bits = (tw * th).bit_length()
    tile = (isect_ids >> 32) & ((1 << tile_bits) - 1)
    tyy, txx = torch.div(tile, tw, rounding_mode="floor"), tile % tw
    keep = (txx >= tx0) & (txx < tx1) & (tyy >= ty0) & (tyy < ty1)
    return isect_ids[keep], flatten_ids[keep]


def compare_window(ref: Dict, idx: torch.Tensor, got_render, got_alpha, got_ids, got_flat, got_grads: Dict, W: int, H: int,
                   window, tile: int = 16) -> Dict:
    """Parity report of the CUDA path's full-frame result (`got_*`, CPU tensors; ids / flat = its full sorted list for
    the one camera, grads = gradients of the FULL scene's leaves) against the oracle window result `ref`."""
    tx0, ty0, tx1, ty1 = window
    tw, th = math.ceil(W / tile), math.ceil(H / tile)
    ys, xs = slice(ty0 * tile, min(H, ty1 * tile)), slice(tx0 * tile, min(W, tx1 * tile))
    gi, gf = window_key_slice(got_ids, got_flat, tw, th, window)
    keys_equal = bool(gi.numel() == ref["isect_ids"].numel() and torch.equal(gi, ref["isect_ids"]))
    ids_equal = bool(gf.numel() == ref["flatten_ids"].numel() and torch.equal(gf.long(), idx[ref["flatten_ids"].long()]))
    fr = ref["fragile"][:, ys, xs]
    dr = (got_render[:, ys, xs] - ref["render"][:, ys, xs].float()).abs().amax(dim=-1)
    da = (got_alpha[:, ys, xs] - ref["alpha"][:, ys, xs].float()).abs().amax(dim=-1)
    ok = ~fr
    rep = {
        "window_tiles": list(window), "window_pixels": int(fr.numel()), "n_gaussians": int(idx.numel()),
        "n_keys": int(ref["isect_ids"].numel()), "keys_bit_equal": keys_equal, "ids_bit_equal": ids_equal,
        "fragile_frac": float(fr.float().mean()),
        "render_max_err": float(dr[ok].max()) if ok.any() else 0.0, "alpha_max_err": float(da[ok].max()) if ok.any() else 0.0,
        "render_max_err_on_fragile": float(dr[fr].max()) if fr.any() else 0.0,
        "alpha_max_err_on_fragile": float(da[fr].max()) if fr.any() else 0.0,
    }
    grel = {}
    outside = torch.ones(got_grads["means"].shape[0], dtype=torch.bool)
    outside[idx] = False
    leak = 0.0
    for k in LEAVES:
        g, r = got_grads[k].double(), ref["grads"][k].double()
        grel[k] = float((g[idx] - r).norm() / r.norm().clamp_min(1e-30))
        leak = max(leak, float(g[outside].abs().max()) if outside.any() else 0.0)
    rep["grad_rel_err"] = grel
    rep["grad_outside_selection_max"] = leak
    rep["pass"] = bool(keys_equal and ids_equal and rep["render_max_err"] <= 1e-4 and rep["alpha_max_err"] <= 1e-4
                       and max(grel.values()) < 1e-3 and leak == 0.0)
    return rep
