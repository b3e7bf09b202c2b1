"""``rasterization()`` -- drop-in for ``gsplat.rendering.rasterization``.

This is the call the reference's splat-training stage makes on every iteration: nerfstudio
``SplatfactoModel.get_outputs`` behind ``ns-train splatfacto`` (source/container/src/main.py:1308-1316,
``packed=False``, ``render_mode="RGB+ED"``) and ``simple_trainer.Runner.rasterize_splats`` behind
``simple_trainer.py --packed`` (source/container/src/main.py:1339-1347, ``packed=True``).

Same argument surface and returned ``(render_colors, render_alphas, meta)``.  Every stage runs in a
hand-written sm_100a kernel of ``libb2s.so``; there is no CPU path and no other backend, and
options that are not implemented raise ``NotImplementedError`` instead of falling back.
"""
from __future__ import annotations

import weakref

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

import os

from . import ops
from ._lib import require_cuda


def _use_coarse_bins() -> bool:
    """Two-level binning (coarse bins expanded inside the compositing kernels) is the default;
    B2S_BINNING=tiles selects the reference-shaped pipeline (materialise every tile intersection) for A/B runs."""
    return os.environ.get("B2S_BINNING", "coarse") != "tiles"


def rasterization(
    means: Tensor,        # [N,3]
    quats: Tensor,        # [N,4] wxyz, un-normalised
    scales: Tensor,       # [N,3] post-exp
    opacities: Tensor,    # [N] post-sigmoid
    colors: Tensor,       # [N,D] | [C,N,D] | SH: [N,K,3] | [C,N,K,3]
    viewmats: Tensor,     # [C,4,4] world->camera
    Ks: Tensor,           # [C,3,3]
    width: int,
    height: int,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    eps2d: float = 0.3,
    sh_degree: Optional[int] = None,
    packed: bool = True,
    tile_size: int = 16,
    backgrounds: Optional[Tensor] = None,
    render_mode: str = "RGB",
    sparse_grad: bool = False,
    absgrad: bool = False,
    rasterize_mode: str = "classic",
    channel_chunk: int = 32,
    distributed: bool = False,
    camera_model: str = "pinhole",
    covars: Optional[Tensor] = None,
    tile_stripes: bool = False,
    stripe_rows=None,
    **unsupported,
) -> Tuple[Tensor, Tensor, Dict]:
    if unsupported:
        raise NotImplementedError(f"unsupported rasterization() options: {sorted(unsupported)}")
    if camera_model != "pinhole":
        raise NotImplementedError("only camera_model='pinhole' is implemented")
    if covars is not None:
        raise NotImplementedError("precomputed covars are not implemented (pass quats and scales)")
    if sparse_grad:
        raise NotImplementedError("sparse_grad is not implemented")
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16")
    if render_mode not in ("RGB", "D", "ED", "RGB+D", "RGB+ED"):
        raise ValueError(f"unknown render_mode {render_mode!r}")
    if rasterize_mode not in ("classic", "antialiased"):
        raise ValueError(f"unknown rasterize_mode {rasterize_mode!r}")
    if distributed:
        from .distributed import rasterization_distributed
        return rasterization_distributed(
            means, quats, scales, opacities, colors, viewmats, Ks, width, height, near_plane=near_plane,
            far_plane=far_plane, radius_clip=radius_clip, eps2d=eps2d, sh_degree=sh_degree, packed=packed,
            backgrounds=backgrounds, render_mode=render_mode, absgrad=absgrad, rasterize_mode=rasterize_mode,
            tile_stripes=tile_stripes, stripe_rows=stripe_rows)
    if tile_stripes:
        raise ValueError("tile_stripes=True needs distributed=True")

    require_cuda(means, quats, scales, opacities, colors, viewmats, Ks)
    N, C = means.shape[0], viewmats.shape[0]
    if means.shape != (N, 3) or quats.shape != (N, 4) or scales.shape != (N, 3) or opacities.shape != (N,):
        raise ValueError("means [N,3], quats [N,4], scales [N,3], opacities [N] expected")
    if viewmats.shape != (C, 4, 4) or Ks.shape != (C, 3, 3):
        raise ValueError("viewmats [C,4,4] and Ks [C,3,3] expected")
    if sh_degree is None:
        if not ((colors.dim() == 2 and colors.shape[0] == N) or (colors.dim() == 3 and colors.shape[:2] == (C, N))):
            raise ValueError("colors must be [N,D] or [C,N,D] when sh_degree is None")
    else:
        if not ((colors.dim() == 3 and colors.shape[0] == N and colors.shape[2] == 3)
                or (colors.dim() == 4 and colors.shape[:2] == (C, N) and colors.shape[3] == 3)):
            raise ValueError("colors must be [N,K,3] or [C,N,K,3] when sh_degree is set")
        if (sh_degree + 1) ** 2 > colors.shape[-2]:
            raise ValueError("sh_degree needs more SH bases than colors provides")
    if backgrounds is not None and backgrounds.shape[0] != C:
        raise ValueError("backgrounds must be [C,D]")

    antialiased = rasterize_mode == "antialiased"
    width, height = int(width), int(height)
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)

    # ---- stage 1: projection (+ SH colour + tile counts) in one kernel ------------------------------------
    fused_sh = sh_degree is not None and colors.dim() == 3
    campos = None          # the kernels derive the camera centres from viewmats
    radii, means2d, depths, conics, _comps, opac, sh_cols, tiles, dkeys, rowids, total = ops._ProjectFused.apply(
        means, quats, scales, opacities, colors if fused_sh else None, viewmats, Ks, campos, width, height,
        float(eps2d), float(near_plane), float(far_plane), float(radius_clip), antialiased, False,
        int(sh_degree) if fused_sh else 0, tile_size)

    if sh_degree is None:
        cols = colors if colors.dim() == 3 else colors[None].expand(C, N, colors.shape[-1])
    elif fused_sh:
        cols = sh_cols
    else:  # per-camera coefficients [C,N,K,3]
        campos = torch.linalg.inv(viewmats.float())[:, :3, 3]
        dirs = means[None, :, :] - campos[:, None, :]
        cols = ops.spherical_harmonics(sh_degree, dirs, colors, masks=radii > 0)
        cols = torch.clamp_min(cols + 0.5, 0.0)

    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros_like(backgrounds[:, :1])], dim=-1)
    elif render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros_like(backgrounds[:, :1])

    # ---- packed layout ----------------------------------------------------------------------------------------
    # The pipeline always runs on the dense [C,N] rows (the kernels address rows by index, so nothing is gained by
    # compacting 92 % of them first); with packed=True the packed tensors of `meta` (camera_ids, gaussian_ids, radii,
    # means2d, ...: the nnz visible pairs, camera-major / Gaussian-ascending) are produced on first access by one
    # selection kernel and one gather kernel, see _packed_meta.
    camera_ids = gaussian_ids = None
    radii_m, means2d_m, depths_m, conics_m, opac_m, cols_m, tiles_m = radii, means2d, depths, conics, opac, cols, tiles

    # ---- stage 2: binning --------------------------------------------------------------------------------
    tiles_per_gauss = tiles_m
    ch = cols_m.shape[-1]
    fused_keys = C <= ops.MAX_CAMS
    if _use_coarse_bins() and ops.coarse_bins_supported(C, tw, th, ch):
        # two-level binning: coarse bins now, per-tile lists expanded inside the compositing kernels; the reference's
        # per-intersection lists (flatten_ids, isect_offsets, isect_ids) are built only if somebody reads them
        bins = ops.bin_coarse(means2d_m, radii_m, depths_m, tile_size, tw, th, C, camera_ids=camera_ids,
                              depth_keys=dkeys if fused_keys else None, row_ids=rowids if fused_keys else None, totals=total)
        flatten_ids = isect_offsets = None
        cache = {}

        def lists():
            if not cache:
                flat, off, mk = ops.bin_depth_first(means2d_m.detach(), radii_m, depths_m.detach(), tiles_m, tile_size, tw, th,
                                                    C, camera_ids=camera_ids, total_isects=total, sorted_ids=bins.sorted_ids)
                cache.update(flat=flat, off=off, mk=mk)
            return cache
    else:
        bins = None
        flatten_ids, isect_offsets, make_isect_ids = ops.bin_depth_first(
            means2d_m, radii_m, depths_m, tiles_m, tile_size, tw, th, C, camera_ids=camera_ids,
            depth_keys=dkeys if fused_keys else None, row_ids=rowids if fused_keys else None, total_isects=total)

    # ---- stage 3: compositing ----------------------------------------------------------------------------
    if ch > channel_chunk:
        outs = []
        for c0 in range(0, ch, channel_chunk):
            bg = backgrounds[..., c0:c0 + channel_chunk] if backgrounds is not None else None
            rc, ra = ops.rasterize_to_pixels(means2d_m, conics_m, cols_m[..., c0:c0 + channel_chunk], opac_m, width, height,
                                             tile_size, isect_offsets, flatten_ids, backgrounds=bg, packed=False,
                                             absgrad=absgrad, bins=bins, n_cameras=C)
            outs.append(rc)
        render_colors, render_alphas = torch.cat(outs, dim=-1), ra
    else:
        render_colors, render_alphas = ops.rasterize_to_pixels(
            means2d_m, conics_m, cols_m, opac_m, width, height, tile_size, isect_offsets, flatten_ids,
            backgrounds=backgrounds, packed=False, absgrad=absgrad, bins=bins, n_cameras=C)
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=1e-10)], dim=-1)

    meta = ops.LazyMeta({
        "camera_ids": camera_ids, "gaussian_ids": gaussian_ids, "radii": radii_m, "means2d": means2d_m,
        "depths": depths_m, "conics": conics_m, "opacities": opac_m, "tile_width": tw, "tile_height": th,
        "tiles_per_gauss": tiles_per_gauss, "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    })
    if bins is not None:
        meta["n_isects"], meta["n_coarse_records"] = bins.n_isects, bins.n_records
        meta.set_lazy("flatten_ids", lambda: lists()["flat"])
        meta.set_lazy("isect_offsets", lambda: lists()["off"])
        meta.set_lazy("isect_ids", lambda: lists()["mk"]())
    else:
        meta["flatten_ids"], meta["isect_offsets"] = flatten_ids, isect_offsets
        meta.set_lazy("isect_ids", make_isect_ids)     # 64-bit sorted keys, rebuilt on first access
    if packed:
        meta = _packed_meta(meta, N, C, cols_m, bins.n_visible if bins is not None else None)
    return render_colors, render_alphas, meta


def _packed_meta(dense: "ops.LazyMeta", N: int, C: int, cols: Tensor, nnz: Optional[int]) -> "ops.LazyMeta":
    """``meta`` of ``rasterization(packed=True)``: the per-pair tensors restricted to the nnz visible (camera, Gaussian)
    pairs, camera-major / Gaussian-ascending, ``camera_ids`` / ``gaussian_ids`` int64 and ``flatten_ids`` re-based on the
    packed rows -- built on first access from the dense rows the pipeline ran on (one selection kernel, one gather kernel;
    the gather is differentiable).  ``meta["means2d"]`` receives ``.grad`` (and ``.absgrad``) after the backward, as the
    densification strategies of the reference's trainer expect (``retain_grad()`` on it is honoured): a hook on the dense
    tensor gathers its gradient into the packed one."""
    cache = {}

    def pk():
        if not cache:
            radii = dense["radii"]
            sel = ops.select_visible(radii, nnz)
            d_m2 = dense["means2d"]
            cam, gid, r_p, m2_p, d_p, c_p, o_p, col_p, t_p = ops.pack_visible(radii, d_m2, dense["depths"], dense["conics"],
                                                                               dense["opacities"], cols, dense["tiles_per_gauss"], sel=sel)
            if d_m2.requires_grad:
                # weak references: the hook lives on the autograd node of the dense tensor, which the packed tensor's own
                # graph points to -- a closure holding either tensor strongly closes a cycle through the C++ graph, and
                # every step's 340 MB of packed rows then stays alive until the cyclic collector runs (measured: reserved
                # memory 3.5 -> 17 GB in 60 steps and a cudaMalloc stall of up to 200 ms every few steps)
                packed_ref, dense_ref = weakref.ref(m2_p), weakref.ref(d_m2)

                def hook(g):
                    m_p, d = packed_ref(), dense_ref()
                    if m_p is None:
                        return
                    m_p.grad = ops.gather_rows2(g, sel)
                    ab = getattr(d, "absgrad", None) if d is not None else None
                    if ab is not None:
                        m_p.absgrad = ops.gather_rows2(ab, sel)
                d_m2.register_hook(hook)
            cache.update(sel=sel, camera_ids=cam, gaussian_ids=gid, radii=r_p, means2d=m2_p, depths=d_p, conics=c_p,
                         opacities=o_p, tiles_per_gauss=t_p)
        return cache

    def flat():
        sel = pk()["sel"]
        inv = torch.full((C * N,), -1, dtype=torch.int32, device=sel.device)
        inv[sel] = torch.arange(sel.numel(), dtype=torch.int32, device=sel.device)
        return inv[dense["flatten_ids"].long()]

    meta = ops.LazyMeta({k: dense[k] for k in ("tile_width", "tile_height", "width", "height", "tile_size", "n_cameras")})
    for k in ("n_isects", "n_coarse_records"):
        if k in dense:
            meta[k] = dense[k]
    for k in ("camera_ids", "gaussian_ids", "radii", "means2d", "depths", "conics", "opacities", "tiles_per_gauss"):
        meta.set_lazy(k, (lambda kk: (lambda: pk()[kk]))(k))
    meta.set_lazy("flatten_ids", flat)
    meta.set_lazy("isect_offsets", lambda: dense["isect_offsets"])
    meta.set_lazy("isect_ids", lambda: dense["isect_ids"])
    return meta
