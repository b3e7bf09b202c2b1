"""CPU oracle for the ``rasterization()`` hot path  --  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product package
(``pipeline_pointcloud_b200``) never imports anything from ``oracle/`` and has
no CPU path at all.

PARITY UNPINNED.  The arithmetic on this path lives in a third-party module,
gsplat (nerfstudio-project/gsplat), which the reference installs *unpinned from
the HEAD of main* (``source/container/Dockerfile:220-229``) and which is absent
from ``/root/reference`` and from this container.  The reference's own tests
hold no golden vector for the path (``scripts/test_resolver.py`` tests path
resolution only).  This oracle is therefore a from-spec restatement of the
semantics of gsplat 1.4 ``gsplat/cuda/_torch_impl.py`` + ``gsplat/rendering.py``
as written down in SURVEY.md section 8a (rows 1-10) and Appendix A, anchored on
the reference call sites that make it run:

* ``source/container/src/main.py:1308-1316``  (``ns-train splatfacto``)
* ``source/container/src/main.py:1339-1347``  (``gsplat/examples/simple_trainer.py --packed``)
* ``source/container/src/pipeline/post_processing/gsplat_pt_to_ply.py:50-73``
  (parameter names / layouts: means, quats wxyz, scales, opacities, sh0, shN)

Frozen ambiguities:

1. radius rule = gsplat 1.4: scalar ``ceil(3 * sqrt(lambda_max))``, ``radii`` int32 ``[C, N]``.
2. ``det`` of the blurred 2-D covariance: culled (radius 0) when ``det <= 0``;
   clamped to ``>= 1e-10`` before the divisions.  ``near < z < far`` strict.
3. sort is stable: ties in (camera, tile, depth bits) keep ascending flat index.
4. ``exp`` in alpha is ``torch.exp``; kernels may use ``ex2.approx``.  Pixels whose
   compositing decision sits on a knife edge (alpha == 1/255, T == 1e-4) are
   reported by ``rasterize_to_pixels(..., return_fragile=True)`` so a parity test
   can exclude them instead of loosening its tolerance.
5. packed rows are camera-major, Gaussian-ascending; ids int64.
6. every expression on the *binning* path (radii, means2d, depths -> tile
   rectangles and sort keys) is written as a fixed sequence of single IEEE-754
   fp32 operations (one torch elementwise op = one rounding; no matmul, no
   einsum, no fused multiply-add), so that a kernel using the same sequence with
   ``__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn`` reproduces it bit for bit.  The
   sequence is the "canonical order" documented in DESIGN.md.

Everything is differentiable through ``torch.autograd`` (the masks are treated
as constants, exactly as a hand-written backward does).  ``dtype=torch.float64``
inputs run the same code in double precision.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

ALPHA_MAX = 0.999          # alpha clamp                     
ALPHA_MIN = 1.0 / 255.0    # contribution threshold          
T_STOP = 1e-4              # transmittance early-out         

SH_C0 = 0.2820947917738781
SH_C1 = 0.48860251190292
SH_C2 = (1.092548430592079, 0.9461746957575601, 0.3153915652525201, 0.5462742152960395)
SH_C3 = (0.5900435899266435, 2.890611442640554, 2.285228997322329, 0.4570457994644658,
         1.865881662950577, 1.119528997770346, 1.445305721320277)


def _sqrt(x: Tensor) -> Tensor:
    """Correctly rounded fp32 square root.

    ``torch.sqrt`` on CPU is *not* correctly rounded for fp32 on every ISA (measured here:
    0.7 % of inputs differ by 1 ulp from IEEE ``sqrtss`` with the AVX-512 kernels), which
    would make the canonical order machine-dependent.  sqrt in fp64 followed by one rounding
    to fp32 is correctly rounded (53 >= 2*24+2 bits), and stays differentiable.
    """
    if x.dtype == torch.float32:
        return torch.sqrt(x.double()).to(torch.float32)
    return torch.sqrt(x)


# --------------------------------------------------------------------------------------
# rows 1-2: quaternion/scale -> covariance, world -> camera, perspective EWA projection
# --------------------------------------------------------------------------------------
def quat_to_rotmat(quats: Tensor) -> Tuple[Tensor, ...]:
    """Rotation matrix entries (R00..R22) of un-normalised (w,x,y,z) quaternions.

    Follows gsplat ``_quat_to_rotmat``.  Canonical order:
    n2 = ((w*w + x*x) + y*y) + z*z ; q /= max(sqrt(n2), 1e-12).
    """
    w, x, y, z = quats.unbind(-1)
    n2 = ((w * w + x * x) + y * y) + z * z
    n = _sqrt(n2).clamp_min(1e-12)
    w, x, y, z = w / n, x / n, y / n, z / n
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    r00 = 1.0 - 2.0 * (y2 + z2)
    r01 = 2.0 * (xy - wz)
    r02 = 2.0 * (xz + wy)
    r10 = 2.0 * (xy + wz)
    r11 = 1.0 - 2.0 * (x2 + z2)
    r12 = 2.0 * (yz - wx)
    r20 = 2.0 * (xz - wy)
    r21 = 2.0 * (yz + wx)
    r22 = 1.0 - 2.0 * (x2 + y2)
    return r00, r01, r02, r10, r11, r12, r20, r21, r22


def quat_scale_to_covar(quats: Tensor, scales: Tensor) -> Tuple[Tensor, ...]:
    """Upper triangle (c00,c01,c02,c11,c12,c22) of Sigma = (R S)(R S)^T.

    gsplat ``_quat_scale_to_covar_preci``.
    """
    r00, r01, r02, r10, r11, r12, r20, r21, r22 = quat_to_rotmat(quats)
    sx, sy, sz = scales.unbind(-1)
    m00, m01, m02 = r00 * sx, r01 * sy, r02 * sz
    m10, m11, m12 = r10 * sx, r11 * sy, r12 * sz
    m20, m21, m22 = r20 * sx, r21 * sy, r22 * sz
    c00 = (m00 * m00 + m01 * m01) + m02 * m02
    c01 = (m00 * m10 + m01 * m11) + m02 * m12
    c02 = (m00 * m20 + m01 * m21) + m02 * m22
    c11 = (m10 * m10 + m11 * m11) + m12 * m12
    c12 = (m10 * m20 + m11 * m21) + m12 * m22
    c22 = (m20 * m20 + m21 * m21) + m22 * m22
    return c00, c01, c02, c11, c12, c22


def quat_scale_to_covar_preci(quats: Tensor, scales: Tensor, compute_covar: bool = True,
                              compute_preci: bool = True, triu: bool = False):
    """Full-matrix form of the above, plus the precision matrix R S^-2 R^T."""
    r = torch.stack(quat_to_rotmat(quats), dim=-1).reshape(quats.shape[:-1] + (3, 3))
    covars = precis = None
    if compute_covar:
        m = r * scales[..., None, :]
        covars = m @ m.transpose(-1, -2)
        if triu:
            covars = covars.reshape(covars.shape[:-2] + (9,))
            covars = (covars[..., [0, 1, 2, 4, 5, 8]] + covars[..., [0, 3, 6, 4, 7, 8]]) / 2.0
    if compute_preci:
        p = r * (1.0 / scales[..., None, :])
        precis = p @ p.transpose(-1, -2)
        if triu:
            precis = precis.reshape(precis.shape[:-2] + (9,))
            precis = (precis[..., [0, 1, 2, 4, 5, 8]] + precis[..., [0, 3, 6, 4, 7, 8]]) / 2.0
    return covars, precis


def fully_fused_projection(
    means: Tensor,      # [N,3]
    quats: Tensor,      # [N,4] wxyz, un-normalised
    scales: Tensor,     # [N,3] post-exp
    viewmats: Tensor,   # [C,4,4] world->camera
    Ks: Tensor,         # [C,3,3]
    width: int,
    height: int,
    eps2d: float = 0.3,
    near_plane: float = 0.01,
    far_plane: float = 1e10,
    radius_clip: float = 0.0,
    calc_compensations: bool = False,
):
    """EWA projection of every (camera, Gaussian) pair  (SURVEY 8a row 2).

    Restates gsplat ``_world_to_cam`` + ``_persp_proj`` + ``_fully_fused_projection``
    with the gsplat-1.4 radius rule.  Returns ``radii`` int32 [C,N], ``means2d``
    [C,N,2], ``depths`` [C,N], ``conics`` [C,N,3], ``compensations`` [C,N] or None.
    Rows with ``radii == 0`` are zeroed.
    """
    dt = means.dtype
    C = viewmats.shape[0]
    c00, c01, c02, c11, c12, c22 = (t[None, :] for t in quat_scale_to_covar(quats, scales))
    mx, my, mz = (means[None, :, i] for i in range(3))
    V = viewmats.to(dt)
    v = [[V[:, i, j][:, None] for j in range(4)] for i in range(3)]
    # camera-space mean  t_i = ((R_i0*mx + R_i1*my) + R_i2*mz) + t_i
    tx = ((v[0][0] * mx + v[0][1] * my) + v[0][2] * mz) + v[0][3]
    ty = ((v[1][0] * mx + v[1][1] * my) + v[1][2] * mz) + v[1][3]
    tz = ((v[2][0] * mx + v[2][1] * my) + v[2][2] * mz) + v[2][3]
    # camera-space covariance  A = R Sigma ; Sc = A R^T
    S = [[c00, c01, c02], [c01, c11, c12], [c02, c12, c22]]
    A = [[(v[i][0] * S[0][j] + v[i][1] * S[1][j]) + v[i][2] * S[2][j] for j in range(3)] for i in range(3)]

    def sc(i, j):
        return (A[i][0] * v[j][0] + A[i][1] * v[j][1]) + A[i][2] * v[j][2]

    s00, s01, s02, s11, s12, s22 = sc(0, 0), sc(0, 1), sc(0, 2), sc(1, 1), sc(1, 2), sc(2, 2)

    Kc = Ks.to(dt)
    fx, fy = Kc[:, 0, 0][:, None], Kc[:, 1, 1][:, None]
    cx, cy = Kc[:, 0, 2][:, None], Kc[:, 1, 2][:, None]
    Wf = torch.tensor(float(width), dtype=dt)
    Hf = torch.tensor(float(height), dtype=dt)
    tan_fovx = (0.5 * Wf) / fx
    tan_fovy = (0.5 * Hf) / fy
    lim_xp = (Wf - cx) / fx + 0.3 * tan_fovx
    lim_xn = cx / fx + 0.3 * tan_fovx
    lim_yp = (Hf - cy) / fy + 0.3 * tan_fovy
    lim_yn = cy / fy + 0.3 * tan_fovy

    depth_ok = (tz > near_plane) & (tz < far_plane)
    tzs = torch.where(depth_ok, tz, torch.ones_like(tz))  # keep culled rows finite
    txz = tx / tzs
    tyz = ty / tzs
    txc = tzs * torch.minimum(lim_xp, torch.maximum(-lim_xn, txz))
    tyc = tzs * torch.minimum(lim_yp, torch.maximum(-lim_yn, tyz))
    tz2 = tzs * tzs
    j00 = fx / tzs
    j11 = fy / tzs
    j02 = -((fx * txc) / tz2)
    j12 = -((fy * tyc) / tz2)
    # B = J Sc (2x3), cov2d = B J^T
    b00 = j00 * s00 + j02 * s02
    b01 = j00 * s01 + j02 * s12
    b02 = j00 * s02 + j02 * s22
    b11 = j11 * s11 + j12 * s12
    b12 = j11 * s12 + j12 * s22
    q00 = b00 * j00 + b02 * j02
    q01 = b01 * j11 + b02 * j12
    q11 = b11 * j11 + b12 * j12
    x2d = fx * txz + cx
    y2d = fy * tyz + cy

    det_orig = q00 * q11 - q01 * q01
    a = q00 + eps2d
    d = q11 + eps2d
    det = a * d - q01 * q01
    det_ok = det > 0
    detc = det.clamp_min(1e-10)
    conic_a = d / detc
    conic_b = -(q01 / detc)
    conic_c = a / detc
    comp = _sqrt((det_orig / detc).clamp_min(0.0))
    mid = 0.5 * (a + d)
    lam = mid + _sqrt((mid * mid - det).clamp_min(0.01))
    radius = torch.ceil(3.0 * _sqrt(lam))
    valid = depth_ok & det_ok & (radius > radius_clip)
    inside = (x2d + radius > 0) & (x2d - radius < Wf) & (y2d + radius > 0) & (y2d - radius < Hf)
    keep = valid & inside
    radii = torch.where(keep, radius, torch.zeros_like(radius)).to(torch.int32)
    z = torch.zeros_like(x2d)
    means2d = torch.stack([torch.where(keep, x2d, z), torch.where(keep, y2d, z)], dim=-1)
    depths = torch.where(keep, tz, z)
    conics = torch.stack([torch.where(keep, conic_a, z), torch.where(keep, conic_b, z),
                          torch.where(keep, conic_c, z)], dim=-1)
    compensations = torch.where(keep, comp, z) if calc_compensations else None
    assert radii.shape == (C, means.shape[0])
    return radii, means2d, depths, conics, compensations


# --------------------------------------------------------------------------------------
# row 3: spherical harmonics
# --------------------------------------------------------------------------------------
def eval_sh_bases(degree: int, dirs: Tensor) -> Tensor:
    """Real SH basis values [..., (degree+1)^2] of unit directions."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, SH_C0)]
    if degree >= 1:
        out += [-SH_C1 * y, SH_C1 * z, -SH_C1 * x]
    if degree >= 2:
        xx, yy, zz = x * x, y * y, z * z
        out += [SH_C2[0] * x * y, -SH_C2[0] * y * z, SH_C2[1] * zz - SH_C2[2],
                -SH_C2[0] * x * z, SH_C2[3] * (xx - yy)]
    if degree >= 3:
        out += [-SH_C3[0] * y * (3.0 * xx - yy), SH_C3[1] * x * y * z,
                y * (-SH_C3[2] * zz + SH_C3[3]), z * (SH_C3[4] * zz - SH_C3[5]),
                x * (-SH_C3[2] * zz + SH_C3[3]), SH_C3[6] * z * (xx - yy),
                -SH_C3[0] * x * (xx - 3.0 * yy)]
    if degree >= 4:
        raise NotImplementedError("oracle covers SH degree 0..3 (the path's range)")
    return torch.stack(out, dim=-1)


def spherical_harmonics(degree: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """colour[..., 3] = sum_k Y_k(normalise(dirs)) * coeffs[..., k, :]   (SURVEY 8a row 3).

    No +0.5 / clamp here -- ``rasterization`` applies them, as gsplat does.
    """
    n = dirs.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    basis = eval_sh_bases(degree, dirs / n)                     # [..., K]
    K = basis.shape[-1]
    col = (basis[..., None] * coeffs[..., :K, :]).sum(dim=-2)
    if masks is not None:
        col = torch.where(masks[..., None], col, torch.zeros_like(col))
    return col


# --------------------------------------------------------------------------------------
# rows 4-6: tile intersection, sort, offsets  (integer, bit-exact)
# --------------------------------------------------------------------------------------
def tile_rects(means2d: Tensor, radii: Tensor, tile_size: int, tile_width: int, tile_height: int):
    """Per-pair tile rectangle [min, max)."""
    m = means2d.to(torch.float32) / float(tile_size)
    r = (radii.to(torch.float32) / float(tile_size))[..., None]
    tmin = torch.floor(m - r).to(torch.int32)
    tmax = torch.ceil(m + r).to(torch.int32)
    lim = torch.tensor([tile_width, tile_height], dtype=torch.int32)
    zero = torch.zeros_like(lim)
    tmin = torch.minimum(torch.maximum(tmin, zero), lim)
    tmax = torch.minimum(torch.maximum(tmax, zero), lim)
    return tmin, tmax


def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int,
                tile_height: int, sort: bool = True, packed: bool = False, n_cameras: Optional[int] = None,
                camera_ids: Optional[Tensor] = None, gaussian_ids: Optional[Tensor] = None):
    """(tiles_per_gauss int32, isect_ids int64, flatten_ids int32)  (SURVEY 8a rows 4-5).

    Dense inputs are [C,N,...]; packed inputs are [nnz,...] with ``camera_ids``.
    key = cam << (32 + tile_bits) | tile << 32 | bits(depth as fp32).
    """
    if packed:
        nnz = means2d.shape[0]
        C = int(n_cameras)
        cam = camera_ids.to(torch.int64)
    else:
        C, N = means2d.shape[:2]
        cam = torch.arange(C, dtype=torch.int64)[:, None].expand(C, N).reshape(-1)
        means2d, radii, depths = means2d.reshape(C * N, 2), radii.reshape(C * N), depths.reshape(C * N)
    tmin, tmax = tile_rects(means2d, radii, tile_size, tile_width, tile_height)
    span = (tmax - tmin).to(torch.int64)
    tpg = span[:, 0] * span[:, 1] * (radii > 0)
    tile_bits = (tile_width * tile_height).bit_length()
    n_isects = int(tpg.sum())
    flat = torch.repeat_interleave(torch.arange(tpg.numel(), dtype=torch.int64), tpg)
    start = torch.cumsum(tpg, 0) - tpg
    k = torch.arange(n_isects, dtype=torch.int64) - start[flat]       # rank inside the rect
    w = span[flat, 0].clamp_min(1)
    ty = tmin[flat, 1].to(torch.int64) + k // w                        # row-major: y outer, x inner
    tx = tmin[flat, 0].to(torch.int64) + k % w
    tile_id = ty * tile_width + tx
    depth_bits = depths.to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    isect_ids = (cam[flat] << (32 + tile_bits)) | (tile_id << 32) | depth_bits[flat]
    flatten_ids = flat.to(torch.int32)
    if sort:
        isect_ids, order = torch.sort(isect_ids, stable=True)
        flatten_ids = flatten_ids[order]
    tpg = tpg.to(torch.int32)
    if not packed:
        tpg = tpg.reshape(C, N)
    return tpg, isect_ids, flatten_ids


def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """offsets[c,ty,tx] = first sorted index of that (camera, tile); empty tiles take
    the next tile's start."""
    tile_bits = (tile_width * tile_height).bit_length()
    n_tiles = tile_width * tile_height
    cam = torch.arange(n_cameras, dtype=torch.int64)[:, None]
    tile = torch.arange(n_tiles, dtype=torch.int64)[None, :]
    want = ((cam << tile_bits) | tile).reshape(-1)
    have = isect_ids >> 32
    off = torch.searchsorted(have, want, right=False)
    return off.to(torch.int32).reshape(n_cameras, tile_height, tile_width)


# --------------------------------------------------------------------------------------
# row 7-8: per-tile front-to-back compositing (autograd supplies the backward)
# --------------------------------------------------------------------------------------
def rasterize_to_pixels(
    means2d: Tensor,        # [C,N,2] or [nnz,2]
    conics: Tensor,         # [C,N,3] or [nnz,3]
    colors: Tensor,         # [C,N,ch] or [nnz,ch]
    opacities: Tensor,      # [C,N] or [nnz]
    image_width: int,
    image_height: int,
    tile_size: int,
    isect_offsets: Tensor,  # [C,th,tw] int32
    flatten_ids: Tensor,    # [I] int32
    backgrounds: Optional[Tensor] = None,  # [C,ch]
    return_fragile: bool = False,
    return_last_ids: bool = False,
):
    """Front-to-back alpha compositing per 16x16 tile.

    The upstream torch rasterizer needs a CUDA index kernel and nerfacc; this one
    enumerates (pixel, Gaussian) pairs itself, one tile at a time, so it runs on CPU.
    Returns render [C,H,W,ch], alpha [C,H,W,1] (+ fragile mask [C,H,W] bool,
    + last_ids [C,H,W] int32 when asked).
    """
    C, th, tw = isect_offsets.shape
    ch = colors.shape[-1]
    dt = means2d.dtype
    xy = means2d.reshape(-1, 2)
    con = conics.reshape(-1, 3)
    col = colors.reshape(-1, ch)
    opa = opacities.reshape(-1)
    n_isects = flatten_ids.numel()
    offs = isect_offsets.reshape(-1).tolist() + [n_isects]
    render_rows = [[None] * tw for _ in range(C * th)]
    alpha_rows = [[None] * tw for _ in range(C * th)]
    fragile = torch.zeros(C, image_height, image_width, dtype=torch.bool)
    last_ids = torch.zeros(C, image_height, image_width, dtype=torch.int32)
    fids = flatten_ids.to(torch.int64)
    for c in range(C):
        for ty in range(th):
            y0, y1 = ty * tile_size, min((ty + 1) * tile_size, image_height)
            for tx in range(tw):
                x0, x1 = tx * tile_size, min((tx + 1) * tile_size, image_width)
                t = (c * th + ty) * tw + tx
                lo = offs[t]
                hi = offs[t + 1]
                ph, pw = y1 - y0, x1 - x0
                if hi <= lo:
                    rc = torch.zeros(ph, pw, ch, dtype=dt)
                    ra = torch.zeros(ph, pw, 1, dtype=dt)
                    if backgrounds is not None:
                        rc = rc + backgrounds[c].to(dt)
                    render_rows[c * th + ty][tx] = rc
                    alpha_rows[c * th + ty][tx] = ra
                    continue
                ids = fids[lo:hi]
                gxy, gcon, gcol, gop = xy[ids], con[ids], col[ids], opa[ids]
                py, px = torch.meshgrid(torch.arange(y0, y1, dtype=dt) + 0.5,
                                        torch.arange(x0, x1, dtype=dt) + 0.5, indexing="ij")
                px, py = px.reshape(-1, 1), py.reshape(-1, 1)          # [P,1]
                dx = gxy[None, :, 0] - px                               # [P,G]
                dy = gxy[None, :, 1] - py
                sigma = 0.5 * (gcon[None, :, 0] * dx * dx + gcon[None, :, 2] * dy * dy) + gcon[None, :, 1] * dx * dy
                vis = torch.exp(-sigma)
                alpha = torch.clamp_max(gop[None, :] * vis, ALPHA_MAX)
                valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
                a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
                one_m = 1.0 - a_eff
                t_incl = torch.cumprod(one_m, dim=1)
                stop_here = t_incl.detach() <= T_STOP
                stopped = torch.cumsum(stop_here.to(torch.int32), dim=1) > 0   # this and all later
                contrib = valid & ~stopped
                t_excl = torch.cat([torch.ones_like(t_incl[:, :1]), t_incl[:, :-1]], dim=1)
                wgt = torch.where(contrib, a_eff * t_excl, torch.zeros_like(a_eff))
                rc = wgt @ gcol                                         # [P,ch]
                t_fin = torch.where(contrib, one_m, torch.ones_like(one_m)).prod(dim=1, keepdim=True)
                ra = 1.0 - t_fin
                if backgrounds is not None:
                    rc = rc + t_fin * backgrounds[c].to(dt)[None, :]
                render_rows[c * th + ty][tx] = rc.reshape(ph, pw, ch)
                alpha_rows[c * th + ty][tx] = ra.reshape(ph, pw, 1)
                if return_fragile or return_last_ids:
                    with torch.no_grad():
                        live = ~torch.cat([torch.zeros_like(stopped[:, :1]), stopped[:, :-1]], dim=1)  # evaluated pairs
                        near_a = (torch.abs(alpha * 255.0 - 1.0) < 5e-5) & (sigma >= 0) & live
                        near_s = (torch.abs(sigma) < 1e-6) & live
                        near_t = (torch.abs(t_incl / T_STOP - 1.0) < 2e-3) & live & valid
                        fr = (near_a | near_s | near_t).any(dim=1)
                        fragile[c, y0:y1, x0:x1] = fr.reshape(ph, pw)
                        G = ids.numel()
                        idx = torch.arange(G)[None, :].expand_as(contrib)
                        last = torch.where(contrib, idx, torch.full_like(idx, -1)).max(dim=1).values
                        last_ids[c, y0:y1, x0:x1] = (last + lo).clamp_min(0).reshape(ph, pw).to(torch.int32)
    render = torch.stack([torch.cat([torch.cat(render_rows[c * th + ty], dim=1) for ty in range(th)], dim=0)
                          for c in range(C)], dim=0)
    alphas = torch.stack([torch.cat([torch.cat(alpha_rows[c * th + ty], dim=1) for ty in range(th)], dim=0)
                          for c in range(C)], dim=0)
    out = [render, alphas]
    if return_fragile:
        out.append(fragile)
    if return_last_ids:
        out.append(last_ids)
    return tuple(out)


# --------------------------------------------------------------------------------------
# row 10: the rasterization() wrapper
# --------------------------------------------------------------------------------------
def rasterization(
    means: Tensor, quats: Tensor, scales: Tensor, opacities: Tensor, colors: Tensor,
    viewmats: Tensor, Ks: Tensor, width: int, height: int,
    near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0, eps2d: float = 0.3,
    sh_degree: Optional[int] = None, packed: bool = True, tile_size: int = 16,
    backgrounds: Optional[Tensor] = None, render_mode: str = "RGB", sparse_grad: bool = False,
    absgrad: bool = False, rasterize_mode: str = "classic", channel_chunk: int = 32,
    distributed: bool = False, camera_model: str = "pinhole", return_fragile: bool = False,
    tile_window: Optional[Tuple[int, int, int, int]] = None,
) -> Tuple[Tensor, Tensor, Dict]:
    """gsplat ``rasterization()`` restated.

    Inputs are post-activation.  ``colors`` is [N,D] / [C,N,D] when ``sh_degree`` is
    None, else SH coefficients [N,K,3] / [C,N,K,3].  ``packed`` changes only the
    layout of the intermediate tensors in ``meta``, never the image.

    ``tile_window=(tx0, ty0, tx1, ty1)`` is a checker convenience, not reference semantics: projection and binning
    run on the whole frame, then only the tiles [tx0,tx1) x [ty0,ty1) are composited (everything else renders as
    empty) and ``meta`` carries the sorted lists filtered to those tiles -- a full-size scene can then be checked
    against the CUDA path on a sub-frustum at a CPU cost proportional to the window.
    """
    assert camera_model == "pinhole" and not distributed
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    assert rasterize_mode in ("classic", "antialiased")
    N, C = means.shape[0], viewmats.shape[0]
    radii, means2d, depths, conics, comps = fully_fused_projection(
        means, quats, scales, viewmats, Ks, width, height, eps2d, near_plane, far_plane, radius_clip,
        calc_compensations=(rasterize_mode == "antialiased"))
    opac = opacities[None, :].expand(C, N)
    if comps is not None:
        opac = opac * comps
    vis = radii > 0
    if sh_degree is None:
        cols = colors if colors.dim() == 3 else colors[None].expand(C, N, colors.shape[-1])
    else:
        campos = torch.linalg.inv(viewmats.to(means.dtype))[:, :3, 3]
        dirs = means[None, :, :] - campos[:, None, :]
        coeffs = colors if colors.dim() == 4 else colors[None].expand(C, N, colors.shape[-2], 3)
        cols = spherical_harmonics(sh_degree, dirs, coeffs, masks=vis)
        cols = torch.clamp_min(cols + 0.5, 0.0)
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[..., None]], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros_like(backgrounds[:, :1])], dim=-1)
    elif render_mode in ("D", "ED"):
        cols = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros_like(backgrounds[:, :1])

    tw = math.ceil(width / tile_size)
    th = math.ceil(height / tile_size)
    cam_ids = gauss_ids = None
    if packed:
        sel = vis.reshape(-1).nonzero(as_tuple=True)[0]           # camera-major, Gaussian-ascending
        cam_ids, gauss_ids = sel // N, sel % N
        p_radii, p_means2d = radii.reshape(-1)[sel], means2d.reshape(-1, 2)[sel]
        p_depths, p_conics = depths.reshape(-1)[sel], conics.reshape(-1, 3)[sel]
        p_opac, p_cols = opac.reshape(-1)[sel], cols.reshape(-1, cols.shape[-1])[sel]
        tpg, isect_ids, flatten_ids = isect_tiles(p_means2d, p_radii, p_depths, tile_size, tw, th,
                                                  packed=True, n_cameras=C, camera_ids=cam_ids,
                                                  gaussian_ids=gauss_ids)
        r_in = (p_means2d, p_conics, p_cols, p_opac)
        m_radii, m_means2d, m_depths, m_conics, m_opac = p_radii, p_means2d, p_depths, p_conics, p_opac
    else:
        tpg, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th)
        r_in = (means2d, conics, cols, opac)
        m_radii, m_means2d, m_depths, m_conics, m_opac = radii, means2d, depths, conics, opac
    if tile_window is not None:
        tx0, ty0, tx1, ty1 = tile_window
        tile_bits = (tw * th).bit_length()
        tile = (isect_ids >> 32) & ((1 << tile_bits) - 1)
        tyy, txx = torch.div(tile, tw, rounding_mode="floor"), tile % tw
        keep = (txx >= tx0) & (txx < tx1) & (tyy >= ty0) & (tyy < ty1)
        isect_ids, flatten_ids = isect_ids[keep], flatten_ids[keep]      # a subsequence of a sorted list stays sorted
    offsets = isect_offset_encode(isect_ids, C, tw, th)
    out = rasterize_to_pixels(*r_in, width, height, tile_size, offsets, flatten_ids, backgrounds,
                              return_fragile=return_fragile)
    render, alphas = out[0], out[1]
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1], render[..., -1:] / alphas.clamp_min(1e-10)], dim=-1)
    meta = {
        "camera_ids": cam_ids, "gaussian_ids": gauss_ids, "radii": m_radii, "means2d": m_means2d,
        "depths": m_depths, "conics": m_conics, "opacities": m_opac, "tile_width": tw, "tile_height": th,
        "tiles_per_gauss": tpg, "isect_ids": isect_ids, "flatten_ids": flatten_ids, "isect_offsets": offsets,
        "width": width, "height": height, "tile_size": tile_size, "n_cameras": C,
    }
    if return_fragile:
        meta["fragile"] = out[2]
    return render, alphas, meta
