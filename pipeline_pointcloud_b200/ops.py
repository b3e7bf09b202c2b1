"""Stage-level operators of the rasterization path, mirroring gsplat's ``gsplat/cuda/_wrapper.py``
(same names, argument meaning and error behaviour).  Each is a thin
``torch.autograd.Function`` around one C-ABI call of ``libb2s.so``; torch only provides device
memory, streams and autograd bookkeeping.

Reference anchors: these ops run inside the ``rasterization()`` call that the reference's
Train-Stage1 component makes through ``ns-train splatfacto``
(source/container/src/main.py:1308-1316) or ``simple_trainer.py --packed``
(source/container/src/main.py:1339-1347).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from ._lib import B2SError, check, f32c, ptr

MAX_CAMS = 64  # B2S_MAX_CAMERAS_PER_CALL


class StageTimer:
    """Optional per-stage CUDA-event timing on the launching stream (used by bench.py inside its timed
    region; two event records per stage, no synchronisation until ``summary()``)."""

    def __init__(self):
        self.events = []      # (name, start, end)
        self.launches = 0     # kernels of libb2s.so launched (CUB's sort passes are counted separately)

    class _Range:
        def __init__(self, timer, name, launches):
            self.t, self.name, self.launches = timer, name, launches

        def __enter__(self):
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
            return self

        def __exit__(self, *exc):
            self.e.record()
            self.t.events.append((self.name, self.s, self.e))
            self.t.launches += self.launches

    def range(self, name, launches=1):
        return StageTimer._Range(self, name, launches)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, s, e in self.events:
            tot, n = out.get(name, (0.0, 0))
            out[name] = (tot + s.elapsed_time(e), n + 1)
        return {k: {"ms_total": v[0], "calls": v[1], "ms_avg": v[0] / v[1]} for k, v in out.items()}


class _NullRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


TIMER: Optional[StageTimer] = None


def _stage(name, launches=1):
    return TIMER.range(name, launches) if TIMER is not None else _NullRange()


def _bit_length(v: int) -> int:
    return int(v).bit_length()


# --------------------------------------------------------------------------------------
# zero fills on a second stream (OFF by default: a measured negative result, kept switchable with B2S_PREFILL=1).
# The backward accumulates into two large zero-filled buffers: the dense SH gradient (12 K bytes per Gaussian: 1.1 GB at
# 5.8 M, 0.17 ms of pure HBM writes) and the slab of per-row gradients compositing adds into atomically.  Neither
# depends on the forward, and the forward's compositing kernel is issue bound with HBM 96 % idle, so filling them on a
# second stream while it runs looked free.  It is not (profiles/r02_variants_c9.txt, gpurun_out r2c10/r2c11): beside the
# L2-resident depth sort the fill costs the sort 0.15 ms; beside compositing -- even as a thin 2-CTA-per-SM kernel that
# reaches 6.2 TB/s alone -- compositing slows by 0.14 ms (its gathers queue behind 6 TB/s of writes in L2) and the step
# gets 0.16 ms LONGER (2.50 vs 2.34 ms).  The fills therefore stay where round 1 had them, in the backward.
# --------------------------------------------------------------------------------------
_SIDE_STREAMS = {}
_HOST_RINGS = {}           # (device, shape, dtype) -> [four pinned host tensors, next slot]
_PENDING_FILLS = []        # deferred fills: (tensor, holder), launched when the compositing kernel is about to be enqueued


def _prefill_enabled() -> bool:
    return os.environ.get("B2S_PREFILL", "0") == "1"


def _launch_fill(t: Tensor, holder: dict) -> None:
    dev = t.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
    ev0 = torch.cuda.Event()
    ev0.record(main)
    side.wait_event(ev0)                       # earlier users of this block (the caching allocator reuses memory) are done
    check(L.lib().b2s_zero_fill(ptr(t), t.numel() * 4, side.cuda_stream), "b2s_zero_fill")
    if TIMER is not None:
        TIMER.launches += 1
    ev1 = torch.cuda.Event()
    ev1.record(side)
    t.record_stream(side)                      # the block is not handed out again before the fill has run
    holder["ev"] = ev1


def zeros_f32(numel: int, dev) -> Tensor:
    """fp32 zeros[numel] on the current stream, filled by the library's own streaming kernel (b2s_zero_fill)."""
    t = torch.empty(max(int(numel), 1), dtype=torch.float32, device=dev)
    check(L.lib().b2s_zero_fill(ptr(t), t.numel() * 4, L.stream_ptr(dev)), "b2s_zero_fill")
    if TIMER is not None:
        TIMER.launches += 1
    return t[:numel]


def prefill_zeros(numel: int, dev, defer: bool = False):
    """fp32 zeros[numel] filled on the side stream: returns (tensor, holder); ``holder["ev"]`` is the event that marks the
    fill done once it has been launched.  ``defer``: launch it with ``flush_prefills`` -- stage 1 creates its buffer
    early but the fill must not run beside the depth sort (L2 resident: 1.3 GB of streaming writes next to it cost the
    sort 0.15 ms, measured), only beside compositing."""
    t = torch.empty(max(int(numel), 1), dtype=torch.float32, device=dev)       # allocated on the caller's stream
    holder = {"ev": None}
    if defer:
        if len(_PENDING_FILLS) >= 8:           # forwards that never composited: do not hold on to their buffers
            flush_prefills(dev)
        _PENDING_FILLS.append((t, holder))
    else:
        _launch_fill(t, holder)
    return t, holder


def flush_prefills(dev) -> None:
    keep = []
    for t, holder in _PENDING_FILLS:
        if t.device == dev:
            _launch_fill(t, holder)
        else:
            keep.append((t, holder))
    _PENDING_FILLS[:] = keep


def take_prefilled(pre, numel: int, dev) -> Tensor:
    """The zeros of a ``prefill_zeros`` pair, ready on the current stream (a fill that was never launched runs now)."""
    t, holder = pre
    if holder["ev"] is None:
        _PENDING_FILLS[:] = [(x, h) for x, h in _PENDING_FILLS if h is not holder]
        check(L.lib().b2s_zero_fill(ptr(t), t.numel() * 4, L.stream_ptr(dev)), "b2s_zero_fill")
    else:
        torch.cuda.current_stream(dev).wait_event(holder["ev"])
    return t[:numel]


# ======================================================================================
# stage 1: fused projection (+SH colour, + tile count)
# ======================================================================================
class _ProjectFused(torch.autograd.Function):
    """radii, means2d, depths, conics, compensations, opacities_out, colors, tiles_per_gauss, depth_keys, row_ids,
    total_isects."""

    @staticmethod
    def forward(ctx, means, quats, scales, opacities, sh_coeffs, viewmats, Ks, campos, width, height, eps2d,
                near_plane, far_plane, radius_clip, antialiased, calc_compensations, sh_degree, tile_size):
        dev = L.require_cuda(means, quats, scales, viewmats, Ks)
        means, quats, scales = f32c(means), f32c(quats), f32c(scales)
        opacities, sh_coeffs = f32c(opacities), f32c(sh_coeffs)
        viewmats, Ks, campos = f32c(viewmats), f32c(Ks), f32c(campos)
        N, C = means.shape[0], viewmats.shape[0]
        if quats.shape != (N, 4) or scales.shape != (N, 3) or means.shape != (N, 3):
            raise B2SError("means [N,3], quats [N,4], scales [N,3] expected")
        if viewmats.shape != (C, 4, 4) or Ks.shape != (C, 3, 3):
            raise B2SError("viewmats [C,4,4], Ks [C,3,3] expected")
        with_sh = sh_coeffs is not None
        if with_sh and (sh_coeffs.dim() != 3 or sh_coeffs.shape[0] != N or sh_coeffs.shape[2] != 3):
            raise B2SError("sh coefficients must be [N,K,3]")
        o = dict(device=dev)
        radii = torch.empty(C, N, dtype=torch.int32, **o)
        means2d = torch.empty(C, N, 2, dtype=torch.float32, **o)
        depths = torch.empty(C, N, dtype=torch.float32, **o)
        conics = torch.empty(C, N, 3, dtype=torch.float32, **o)
        comps = torch.empty(C, N, dtype=torch.float32, **o) if (calc_compensations or antialiased) else None
        opac_out = torch.empty(C, N, dtype=torch.float32, **o) if opacities is not None else None
        colors = torch.empty(C, N, 3, dtype=torch.float32, **o) if with_sh else None
        tiles = torch.empty(C, N, dtype=torch.int32, **o) if tile_size else None
        dkeys = torch.empty(C, N, dtype=torch.int32, **o) if tile_size else None     # uint32 bit patterns
        rowids = torch.empty(C, N, dtype=torch.int32, **o) if tile_size else None
        total = zeros_f32(32 * 4 * 2, dev).view(torch.int64).view(32, 4) if tile_size else None   # B2S_TOTAL_SLOTS x [tiles, visible rows, coarse records, -]
        tw = (width + tile_size - 1) // tile_size if tile_size else 0
        th = (height + tile_size - 1) // tile_size if tile_size else 0
        lib = L.lib()
        st = L.stream_ptr(dev)
        for c0 in range(0, C, MAX_CAMS):
            c1 = min(C, c0 + MAX_CAMS)
            a = L.ProjectFwdArgs(
                N=N, C=c1 - c0, means=ptr(means), quats=ptr(quats), scales=ptr(scales), opacities=ptr(opacities),
                viewmats=ptr(viewmats[c0:c1]), Ks=ptr(Ks[c0:c1]), width=width, height=height, eps2d=eps2d,
                near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip, antialiased=int(antialiased),
                sh_coeffs=ptr(sh_coeffs), campos=ptr(campos[c0:c1]) if campos is not None else None,
                sh_degree=int(sh_degree) if with_sh else 0, sh_stride_k=sh_coeffs.shape[1] if with_sh else 0,
                tile_size=int(tile_size or 0), tile_w=tw, tile_h=th,
                radii=ptr(radii[c0:c1]), means2d=ptr(means2d[c0:c1]), depths=ptr(depths[c0:c1]),
                conics=ptr(conics[c0:c1]), compensations=ptr(comps[c0:c1]) if comps is not None else None,
                opacities_out=ptr(opac_out[c0:c1]) if opac_out is not None else None,
                colors=ptr(colors[c0:c1]) if colors is not None else None,
                tiles_per_gauss=ptr(tiles[c0:c1]) if tiles is not None else None,
                depth_keys=ptr(dkeys[c0:c1]) if dkeys is not None else None,
                row_ids=ptr(rowids[c0:c1]) if rowids is not None else None, total_isects=ptr(total))
            if rowids is not None and c0 > 0:
                raise NotImplementedError("more than 64 cameras per call with fused binning keys")
            with _stage("project_fwd"):
                check(lib.b2s_project_fwd(ctypes.byref(a), st), "b2s_project_fwd")
        ctx.save_for_backward(means, quats, scales, opacities, sh_coeffs, viewmats, Ks, campos, radii)
        ctx.cfg = (width, height, eps2d, near_plane, far_plane, radius_clip, bool(antialiased), sh_degree)
        ctx.v_sh_pre = None
        if with_sh and ctx.needs_input_grad[4] and C <= MAX_CAMS and N > 0 and _prefill_enabled():
            ctx.v_sh_pre = prefill_zeros(sh_coeffs.numel(), dev, defer=True)   # the dense SH gradient, zero-filled under compositing
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii)
        if tiles is not None:
            ctx.mark_non_differentiable(tiles, dkeys, rowids, total)
        return radii, means2d, depths, conics, comps, opac_out, colors, tiles, dkeys, rowids, total

    @staticmethod
    def backward(ctx, _v_radii, v_means2d, v_depths, v_conics, v_comps, v_opac_out, v_colors, _v_tiles=None, _v_keys=None,
                 _v_rows=None, _v_total=None):
        # viewmats: gsplat's `v_viewmats` (nerfstudio's camera optimizer); Ks: no gradient, as upstream
        want_view = ctx.needs_input_grad[5]
        pre, ctx.v_sh_pre = ctx.v_sh_pre, None          # (a second backward of the same graph gets a fresh buffer)
        grads = project_backward(ctx.saved_tensors, ctx.cfg, v_means2d, v_depths, v_conics, v_comps, v_opac_out, v_colors,
                                 want_viewmats=want_view, v_sh_pre=pre)
        return tuple(grads[:5]) + ((grads[5] if want_view else None),) + (None,) * 12


def project_backward(saved, cfg, v_means2d, v_depths, v_conics, v_comps, v_opac_out, v_colors, want_viewmats=False, v_sh_pre=None):
    """VJP of stage 1 from dense ``[C,N,..]`` cotangents (any of them may be None): (v_means, v_quats, v_scales,
    v_opacities, v_sh_coeffs[, v_viewmats]).  Shared by ``_ProjectFused`` and the multi-GPU ``_ProjectRoute``."""
    means, quats, scales, opacities, sh_coeffs, viewmats, Ks, campos, radii = saved
    width, height, eps2d, near_plane, far_plane, radius_clip, antialiased, sh_degree = cfg
    N, C = means.shape[0], viewmats.shape[0]
    dev = means.device
    v_means2d, v_depths, v_conics = f32c(v_means2d), f32c(v_depths), f32c(v_conics)
    v_comps, v_opac_out, v_colors = f32c(v_comps), f32c(v_opac_out), f32c(v_colors)
    with_sh = sh_coeffs is not None
    v_means = torch.empty_like(means)
    v_quats = torch.empty_like(quats)
    v_scales = torch.empty_like(scales)
    v_opac = torch.empty_like(opacities) if opacities is not None else None
    v_sh, prezeroed = None, 0
    if with_sh:
        K = (sh_degree + 1) ** 2
        if v_sh_pre is not None:              # zero-filled on the side stream during the forward (prefill_zeros)
            v_sh, prezeroed = take_prefilled(v_sh_pre, sh_coeffs.numel(), dev).view(sh_coeffs.shape), 1
        else:
            v_sh = torch.empty_like(sh_coeffs) if K == sh_coeffs.shape[1] else torch.zeros_like(sh_coeffs)
    if N == 0:
        return (v_means, v_quats, v_scales, v_opac, v_sh) + ((torch.zeros_like(viewmats),) if want_viewmats else ())
    lib = L.lib()
    st = L.stream_ptr(dev)
    v_view = torch.zeros(C, 16, device=dev) if want_viewmats else None
    # scratch for the sparse backward (rows with all-zero cotangents are filtered once; include/b2s.h)
    ws = None
    if C <= MAX_CAMS and os.environ.get("B2S_SPARSE_BWD", "1") != "0":
        ws = torch.empty(int(lib.b2s_project_bwd_workspace_bytes(N)), dtype=torch.uint8, device=dev)

    def sl(t, c0, c1):
        return ptr(t[c0:c1]) if t is not None else None

    for c0 in range(0, max(C, 1), MAX_CAMS):
        c1 = min(C, c0 + MAX_CAMS)

        def args(stages):
            return L.ProjectBwdArgs(
                N=N, C=c1 - c0, means=ptr(means), quats=ptr(quats), scales=ptr(scales), opacities=ptr(opacities),
                viewmats=ptr(viewmats[c0:c1]), Ks=ptr(Ks[c0:c1]), width=width, height=height, eps2d=eps2d,
                near_plane=near_plane, far_plane=far_plane, radius_clip=radius_clip, antialiased=int(antialiased),
                sh_coeffs=ptr(sh_coeffs), campos=sl(campos, c0, c1), sh_degree=int(sh_degree) if with_sh else 0,
                sh_stride_k=sh_coeffs.shape[1] if with_sh else 0, radii=ptr(radii[c0:c1]),
                v_means2d=sl(v_means2d, c0, c1), v_depths=sl(v_depths, c0, c1), v_conics=sl(v_conics, c0, c1),
                v_compensations=sl(v_comps, c0, c1), v_opacities_out=sl(v_opac_out, c0, c1),
                v_colors=sl(v_colors, c0, c1), accumulate=int(c0 > 0),
                v_means=ptr(v_means), v_quats=ptr(v_quats), v_scales=ptr(v_scales), v_opacities=ptr(v_opac),
                v_sh_coeffs=ptr(v_sh), stages=stages, workspace=ptr(ws), workspace_bytes=ws.numel() if ws is not None else 0,
                v_sh_prezeroed=prezeroed)

        sparse = ws is not None and c0 == 0      # kernels launched: + classify, zero fill and the two list-driven ones
        if with_sh and TIMER is not None:     # time the stages of the call separately (bench.py)
            if sparse:
                with _stage("bwd_classify", launches=1):
                    check(lib.b2s_project_bwd(ctypes.byref(args(3)), st), "b2s_project_bwd")
                with _stage("sh_bwd", launches=3):
                    check(lib.b2s_project_bwd(ctypes.byref(args(4)), st), "b2s_project_bwd")
            else:
                with _stage("sh_bwd", launches=1):
                    check(lib.b2s_project_bwd(ctypes.byref(args(1)), st), "b2s_project_bwd")
            with _stage("project_bwd", launches=1):
                check(lib.b2s_project_bwd(ctypes.byref(args(2)), st), "b2s_project_bwd")
        else:
            with _stage("project_bwd", launches=(5 if with_sh else 2) if sparse else (2 if with_sh else 1)):
                check(lib.b2s_project_bwd(ctypes.byref(args(0)), st), "b2s_project_bwd")
        if want_viewmats:
            with _stage("viewmat_bwd"):
                check(lib.b2s_viewmat_bwd(ctypes.byref(args(0)), ptr(v_view[c0:c1]), st), "b2s_viewmat_bwd")
    if not want_viewmats:
        return v_means, v_quats, v_scales, v_opac, v_sh
    # assemble [C,4,4]: rotation block and translation column; the SH view directions reach the camera through its
    # centre campos = -A^-1 t (as torch.linalg.inv does upstream): d campos = -A^-1 dA campos - A^-1 dt
    A, t = viewmats[:, :3, :3].double(), viewmats[:, :3, 3].double()
    d_rot, d_t, d_cam = v_view[:, :9].reshape(C, 3, 3).double(), v_view[:, 9:12].double(), v_view[:, 12:15].double()
    if campos is None:
        cam_c = -torch.linalg.solve(A, t.unsqueeze(-1)).squeeze(-1)
        w = torch.linalg.solve(A.transpose(1, 2), d_cam.unsqueeze(-1)).squeeze(-1)
        d_rot = d_rot - w.unsqueeze(-1) * cam_c.unsqueeze(-2)
        d_t = d_t - w
    v_viewmats = torch.zeros(C, 4, 4, dtype=torch.float64, device=dev)
    v_viewmats[:, :3, :3], v_viewmats[:, :3, 3] = d_rot, d_t
    return v_means, v_quats, v_scales, v_opac, v_sh, v_viewmats.to(viewmats.dtype)


def fully_fused_projection(
    means: Tensor, covars: Optional[Tensor], quats: Optional[Tensor], scales: Optional[Tensor], viewmats: Tensor,
    Ks: Tensor, width: int, height: int, eps2d: float = 0.3, near_plane: float = 0.01, far_plane: float = 1e10,
    radius_clip: float = 0.0, packed: bool = False, sparse_grad: bool = False, calc_compensations: bool = False,
    camera_model: str = "pinhole",
):
    """gsplat ``fully_fused_projection``: (radii, means2d, depths, conics, compensations), all ``[C,N,...]``;
    with ``packed=True``: (camera_ids, gaussian_ids, radii, means2d, depths, conics, compensations) over the
    ``nnz`` visible pairs, camera-major / Gaussian-ascending."""
    if camera_model != "pinhole":
        raise NotImplementedError("only camera_model='pinhole' is implemented")
    if covars is not None:
        raise NotImplementedError("pass quats and scales (precomputed covars are not implemented)")
    if sparse_grad:
        raise NotImplementedError("sparse_grad is not implemented")
    if quats is None or scales is None:
        raise ValueError("quats and scales are required")
    radii, means2d, depths, conics, comps = _ProjectFused.apply(
        means, quats, scales, None, None, viewmats, Ks, None, int(width), int(height), float(eps2d), float(near_plane),
        float(far_plane), float(radius_clip), False, bool(calc_compensations), 0, 0)[:5]
    if not packed:
        return radii, means2d, depths, conics, comps
    N = means.shape[0]
    sel = select_visible(radii)
    camera_ids, gaussian_ids = sel // N, sel % N
    return (camera_ids, gaussian_ids, radii.reshape(-1)[sel], means2d.reshape(-1, 2)[sel], depths.reshape(-1)[sel],
            conics.reshape(-1, 3)[sel], comps.reshape(-1)[sel] if comps is not None else None)


@torch.no_grad()
def select_visible(radii: Tensor, nnz: Optional[int] = None) -> Tensor:
    """Indices (int64, ascending: camera-major, Gaussian-ascending) of the rows of dense ``radii`` with radius > 0 -- the
    row order of gsplat's packed mode.  One single-pass scan kernel; ``nnz`` (when the caller already knows it) saves the
    host sync that sizes the result."""
    dev = L.require_cuda(radii)
    radii = radii.contiguous()
    M = radii.numel()
    lib, st = L.lib(), L.stream_ptr(dev)
    cap = M if nnz is None else int(nnz)
    sel = torch.empty(cap, dtype=torch.int64, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.b2s_isect_scan_workspace_bytes(M)) + 16, dtype=torch.uint8, device=dev)
    with _stage("select_visible"):
        check(lib.b2s_select_visible(M, ptr(radii), ptr(sel), cap, ptr(total), ptr(ws), ws.numel(), st), "b2s_select_visible")
    return sel[:int(total.item())] if nnz is None else sel


class _PackVisible(torch.autograd.Function):
    """Dense stage-1 outputs -> packed rows (camera-major, Gaussian-ascending); backward scatters into zeros."""

    @staticmethod
    def forward(ctx, sel, N, means2d, depths, conics, opacities, colors, radii, tiles):
        dev = means2d.device
        nnz, ch = sel.numel(), colors.shape[-1]
        means2d, depths, conics, opacities, colors = f32c(means2d), f32c(depths), f32c(conics), f32c(opacities), f32c(colors)
        o = dict(device=dev)
        means2d_p = torch.empty(nnz, 2, **o)
        depths_p = torch.empty(nnz, **o)
        conics_p = torch.empty(nnz, 3, **o)
        opac_p = torch.empty(nnz, **o)
        colors_p = torch.empty(nnz, ch, **o)
        radii_p = torch.empty(nnz, dtype=torch.int32, **o)
        tiles_p = torch.empty(nnz, dtype=torch.int32, **o) if tiles is not None else None
        cam = torch.empty(nnz, dtype=torch.int64, **o)
        gid = torch.empty(nnz, dtype=torch.int64, **o)
        with _stage("pack_rows"):
            check(L.lib().b2s_pack_rows(nnz, N, ptr(sel), ch, ptr(means2d), ptr(depths), ptr(conics), ptr(opacities),
                                        ptr(colors), ptr(radii), ptr(tiles), ptr(means2d_p), ptr(depths_p), ptr(conics_p),
                                        ptr(opac_p), ptr(colors_p), ptr(radii_p), ptr(tiles_p), ptr(cam), ptr(gid),
                                        L.stream_ptr(dev)), "b2s_pack_rows")
        ctx.save_for_backward(sel)
        ctx.shapes = (means2d.shape, depths.shape, conics.shape, opacities.shape, colors.shape)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(radii_p, cam, gid)
        if tiles_p is not None:
            ctx.mark_non_differentiable(tiles_p)
        return means2d_p, depths_p, conics_p, opac_p, colors_p, radii_p, tiles_p, cam, gid

    @staticmethod
    def backward(ctx, g_means2d, g_depths, g_conics, g_opac, g_colors, *_):
        (sel,) = ctx.saved_tensors
        dev = sel.device
        shapes = ctx.shapes
        gs = [f32c(g) for g in (g_means2d, g_depths, g_conics, g_opac, g_colors)]
        outs = [torch.zeros(s, dtype=torch.float32, device=dev) if g is not None else None for s, g in zip(shapes, gs)]
        ch = shapes[4][-1]
        with _stage("unpack_rows"):
            check(L.lib().b2s_unpack_rows(sel.numel(), ptr(sel), ch, *[ptr(g) for g in gs], *[ptr(o) for o in outs],
                                          L.stream_ptr(dev)), "b2s_unpack_rows")
        return (None, None) + tuple(outs) + (None, None)


def pack_visible(radii: Tensor, means2d: Tensor, depths: Tensor, conics: Tensor, opacities: Tensor, colors: Tensor,
                 tiles: Optional[Tensor] = None, sel: Optional[Tensor] = None):
    """Packed layout of the visible (camera, Gaussian) pairs of dense ``[C,N,..]`` tensors.
    Returns (camera_ids, gaussian_ids, radii, means2d, depths, conics, opacities, colors, tiles)."""
    N = radii.shape[1]
    if sel is None:
        sel = select_visible(radii)
    m2, d, c, o, col, r, t, cam, gid = _PackVisible.apply(sel, N, means2d, depths, conics, opacities, colors,
                                                          radii.contiguous(), tiles.contiguous() if tiles is not None else None)
    return cam, gid, r, m2, d, c, o, col, t


@torch.no_grad()
def gather_rows2(x: Tensor, sel: Tensor) -> Tensor:
    """``x.reshape(-1, 2)[sel]`` with the gather kernel of the packed layout (the gradient of ``meta["means2d"]``)."""
    x = f32c(x)
    out = torch.empty(sel.numel(), 2, device=x.device)
    check(L.lib().b2s_pack_rows(sel.numel(), 1, ptr(sel), 0, ptr(x), None, None, None, None, None, None, ptr(out), None, None, None,
                                None, None, None, None, None, L.stream_ptr(x.device)), "b2s_pack_rows")
    return out


# ======================================================================================
# stand-alone SH / covariance ops
# ======================================================================================
class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree, dirs, coeffs, masks):
        dev = L.require_cuda(dirs, coeffs)
        dirs, coeffs = f32c(dirs), f32c(coeffs)
        K = coeffs.shape[-2]
        if (degree + 1) ** 2 > K:
            raise B2SError("coeffs has fewer than (degree+1)^2 bases")
        if dirs.shape[:-1] != coeffs.shape[:-2] or dirs.shape[-1] != 3 or coeffs.shape[-1] != 3:
            raise B2SError("dirs [...,3] and coeffs [...,K,3] must share their leading shape")
        M = dirs.numel() // 3
        m8 = masks.contiguous().to(torch.uint8) if masks is not None else None
        colors = torch.empty(dirs.shape, dtype=torch.float32, device=dev)
        check(L.lib().b2s_sh_fwd(M, degree, K, ptr(dirs), ptr(coeffs), ptr(m8), ptr(colors), L.stream_ptr(dev)), "b2s_sh_fwd")
        ctx.save_for_backward(dirs, coeffs, m8)
        ctx.degree = degree
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        dirs, coeffs, m8 = ctx.saved_tensors
        degree, K = ctx.degree, coeffs.shape[-2]
        M = dirs.numel() // 3
        v_colors = f32c(v_colors)
        v_coeffs = torch.zeros_like(coeffs) if (degree + 1) ** 2 < K else torch.empty_like(coeffs)
        v_dirs = torch.empty_like(dirs) if ctx.needs_input_grad[1] else None
        check(L.lib().b2s_sh_bwd(M, degree, K, ptr(dirs), ptr(coeffs), ptr(m8), ptr(v_colors), ptr(v_coeffs), ptr(v_dirs),
                                 L.stream_ptr(dirs.device)), "b2s_sh_bwd")
        return None, v_dirs, v_coeffs, None


def spherical_harmonics(degrees_to_use: int, dirs: Tensor, coeffs: Tensor, masks: Optional[Tensor] = None) -> Tensor:
    """gsplat ``spherical_harmonics``: colours [...,3] from dirs [...,3] (normalised inside) and coeffs [...,K,3]."""
    if not 0 <= degrees_to_use <= 3:
        raise NotImplementedError("SH degree 0..3 (the range the training path uses)")
    return _SphericalHarmonics.apply(int(degrees_to_use), dirs, coeffs, masks)


def quat_scale_to_covar_preci(quats: Tensor, scales: Tensor, compute_covar: bool = True, compute_preci: bool = True,
                              triu: bool = False):
    """gsplat ``quat_scale_to_covar_preci`` (forward only: the fused projection never materialises covariances)."""
    if quats.requires_grad or scales.requires_grad:
        raise NotImplementedError("quat_scale_to_covar_preci is forward-only; the training path fuses it into projection")
    dev = L.require_cuda(quats, scales)
    quats, scales = f32c(quats), f32c(scales)
    N = quats.shape[0]
    covars = torch.empty(N, 3, 3, device=dev) if compute_covar else None
    precis = torch.empty(N, 3, 3, device=dev) if compute_preci else None
    check(L.lib().b2s_quat_scale_to_covar_preci(N, ptr(quats), ptr(scales), ptr(covars), ptr(precis), L.stream_ptr(dev)),
          "b2s_quat_scale_to_covar_preci")
    if triu:
        idx = torch.tensor([0, 1, 2, 4, 5, 8], device=dev)
        covars = covars.reshape(N, 9)[:, idx] if covars is not None else None
        precis = precis.reshape(N, 9)[:, idx] if precis is not None else None
    return covars, precis


# ======================================================================================
# stage 2: binning
# ======================================================================================
@torch.no_grad()
def isect_tiles(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
                sort: bool = True, packed: bool = False, n_cameras: Optional[int] = None,
                camera_ids: Optional[Tensor] = None, gaussian_ids: Optional[Tensor] = None,
                tiles_per_gauss: Optional[Tensor] = None, method: str = "radix64",
                tile_rows: Optional[Tuple[int, int]] = None, count_only: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """gsplat ``isect_tiles``: (tiles_per_gauss int32, isect_ids int64, flatten_ids int32), sorted by
    (camera, tile, depth bits) when ``sort``.  ``tiles_per_gauss`` may be passed in when stage 1 already made it.
    ``method="radix64"`` is the reference's algorithm (fill 64-bit keys, LSD radix sort);
    ``method="depth_first"`` is the low-traffic path ``rasterization()`` uses -- same output bit for bit.
    ``tile_rows=(lo, hi)`` restricts the intersections to tile rows [lo, hi) of the grid (a multi-GPU tile stripe):
    the result is the whole-image list filtered by tile row (``tiles_per_gauss`` must then be None or clipped alike)."""
    if method not in ("radix64", "depth_first"):
        raise ValueError("method must be 'radix64' or 'depth_first'")
    dev = L.require_cuda(means2d, radii, depths)
    if packed:
        if camera_ids is None or n_cameras is None:
            raise ValueError("packed=True needs n_cameras and camera_ids")
        M, N, C = means2d.shape[0], 0, int(n_cameras)
        cam = camera_ids.contiguous().to(torch.int64)
        out_shape = (M,)
    else:
        C, N = means2d.shape[:2]
        M, cam = C * N, None
        out_shape = (C, N)
    means2d, depths = f32c(means2d), f32c(depths)
    radii = radii.contiguous().to(torch.int32)
    lib, st = L.lib(), L.stream_ptr(dev)
    if tiles_per_gauss is None:
        tiles_per_gauss = torch.empty(out_shape, dtype=torch.int32, device=dev)
        check(lib.b2s_isect_count(M, ptr(means2d), ptr(radii), tile_size, tile_width, tile_height, ptr(tiles_per_gauss),
                                  L.tile_rows(tile_rows), st), "b2s_isect_count")
    else:
        tiles_per_gauss = tiles_per_gauss.contiguous()
    if count_only:
        return tiles_per_gauss
    if method == "depth_first" and sort:
        flat, _off, make_ids = bin_depth_first(means2d, radii, depths, tiles_per_gauss, tile_size, tile_width, tile_height,
                                               C, camera_ids=cam, tile_rows=tile_rows)
        return tiles_per_gauss, make_ids(), flat
    cum = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.b2s_isect_scan_workspace_bytes(M)), dtype=torch.uint8, device=dev)
    with _stage("isect_scan"):
        check(lib.b2s_isect_scan(M, ptr(tiles_per_gauss), ptr(cum), ptr(total), ptr(ws), ws.numel(), st), "b2s_isect_scan")
    n_isects = int(total.item())          # the one host sync of a forward: sizes the key/value buffers
    if n_isects >= 2 ** 31:
        raise B2SError(f"n_isects = {n_isects} does not fit the int32 offsets of the API")
    isect_ids = torch.empty(n_isects, dtype=torch.int64, device=dev)
    flatten_ids = torch.empty(n_isects, dtype=torch.int32, device=dev)
    with _stage("isect_fill"):
        check(lib.b2s_isect_fill(M, N, ptr(cam), ptr(means2d), ptr(radii), ptr(depths), ptr(cum), tile_size, tile_width,
                                 tile_height, ptr(isect_ids), ptr(flatten_ids), L.tile_rows(tile_rows), st), "b2s_isect_fill")
    if sort and n_isects > 0:
        key_bits = 32 + _bit_length(tile_width * tile_height) + _bit_length(C)
        ids_sorted = torch.empty_like(isect_ids)
        flat_sorted = torch.empty_like(flatten_ids)
        ws = torch.empty(int(lib.b2s_isect_sort_workspace_bytes(n_isects)), dtype=torch.uint8, device=dev)
        with _stage("isect_sort", launches=0):
            check(lib.b2s_isect_sort(n_isects, min(64, key_bits), ptr(isect_ids), ptr(flatten_ids), ptr(ids_sorted),
                                     ptr(flat_sorted), ptr(ws), ws.numel(), st), "b2s_isect_sort")
        isect_ids, flatten_ids = ids_sorted, flat_sorted
    return tiles_per_gauss, isect_ids, flatten_ids


class LazyMeta(dict):
    """The ``meta`` dict of ``rasterization()``.  ``isect_ids`` (the 64-bit sorted keys) is not needed by
    compositing, so the depth-first binning path rebuilds it only when somebody looks at it."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._lazy = {}

    def set_lazy(self, key, fn):
        self._lazy[key] = fn

    def _force(self, key=None):
        for k in ([key] if key is not None else list(self._lazy)):
            if k in self._lazy:
                super().__setitem__(k, self._lazy.pop(k)())

    def __getitem__(self, key):
        self._force(key)
        return super().__getitem__(key)

    def get(self, key, default=None):
        self._force(key)
        return super().get(key, default)

    def __contains__(self, key):
        return key in self._lazy or super().__contains__(key)

    def keys(self):
        self._force()
        return super().keys()

    def items(self):
        self._force()
        return super().items()

    def values(self):
        self._force()
        return super().values()

    def __iter__(self):
        self._force()
        return super().__iter__()

    def __len__(self):
        return super().__len__() + len(self._lazy)


def _async_to_host(t: Tensor):
    """Copies a small device tensor to pinned host memory on the side stream, behind everything enqueued on the current
    stream so far; returns (host tensor, event to ``synchronize()`` before reading it)."""
    dev = t.device
    main = torch.cuda.current_stream(dev)
    side = _SIDE_STREAMS.get(dev.index)
    if side is None:
        side = _SIDE_STREAMS[dev.index] = torch.cuda.Stream(dev)
    produced = torch.cuda.Event()
    produced.record(main)
    # a small ring of persistent pinned buffers per device and shape: the caller reads its buffer before it returns, so a
    # slot is free again long before the ring wraps (no pinned allocation on the per-step path)
    key = (dev.index, tuple(t.shape), t.dtype)
    ring = _HOST_RINGS.get(key)
    if ring is None:
        ring = _HOST_RINGS[key] = [[torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for _ in range(4)], 0]
    host = ring[0][ring[1] % 4]
    ring[1] += 1
    side.wait_event(produced)
    with torch.cuda.stream(side):
        host.copy_(t, non_blocking=True)
    done = torch.cuda.Event()
    done.record(side)
    t.record_stream(side)
    return host, done


def _ws(nbytes: int, dev) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)   # torch allocations are 512-B aligned


@torch.no_grad()
def bin_depth_first(means2d: Tensor, radii: Tensor, depths: Tensor, tiles_per_gauss: Tensor, tile_size: int,
                    tile_width: int, tile_height: int, n_cameras: int, camera_ids: Optional[Tensor] = None,
                    depth_keys: Optional[Tensor] = None, row_ids: Optional[Tensor] = None,
                    total_isects: Optional[Tensor] = None, sorted_ids: Optional[Tensor] = None,
                    tile_rows: Optional[Tuple[int, int]] = None):
    """Depth-first binning (see include/b2s.h): returns (flatten_ids int32 [I], isect_offsets int32 [C,th,tw],
    make_isect_ids) where ``make_isect_ids()`` rebuilds the 64-bit sorted keys.  Bit-identical to
    ``isect_tiles(sort=True)`` + ``isect_offset_encode`` with ~3x less HBM traffic.
    ``depth_keys`` / ``row_ids`` / ``total_isects`` come from stage 1 when it ran fused."""
    dev = L.require_cuda(means2d, radii, depths)
    packed = camera_ids is not None
    M = radii.numel()
    N = 0 if packed else means2d.shape[1]
    C = int(n_cameras)
    means2d, depths = f32c(means2d), f32c(depths)
    radii = radii.contiguous()
    cam = camera_ids.contiguous().to(torch.int64) if packed else None
    lib, st = L.lib(), L.stream_ptr(dev)
    if sorted_ids is None and (depth_keys is None or row_ids is None):
        depth_keys = torch.empty(M, dtype=torch.int32, device=dev)
        row_ids = torch.empty(M, dtype=torch.int32, device=dev)
        with _stage("isect_depth_keys"):
            check(lib.b2s_isect_depth_keys(M, ptr(depths), ptr(radii), ptr(depth_keys), ptr(row_ids), st), "b2s_isect_depth_keys")
    if total_isects is None:
        total_isects = tiles_per_gauss.sum(dtype=torch.int64).reshape(1)
    if sorted_ids is None:
        sorted_ids = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        ws = _ws(lib.b2s_depth_order_workspace_bytes(M), dev)
        with _stage("depth_order"):
            check(lib.b2s_depth_order(M, ptr(depth_keys), ptr(row_ids), ptr(sorted_ids), ptr(ws), ws.numel(), st), "b2s_depth_order")
    # The one host sync of a forward (sizes the per-intersection buffers).  The depth sort is already
    # enqueued, so the GPU keeps working while the host waits for the count that stage 1 produced.
    n_isects = int(sum(r[0] for r in total_isects.reshape(-1, 4).tolist())) if total_isects.numel() > 1 else int(total_isects.item())
    if n_isects >= 2 ** 31:
        raise B2SError(f"n_isects = {n_isects} does not fit the int32 offsets of the API")
    key_bytes = int(lib.b2s_tile_key_bytes(C, tile_width, tile_height))
    tile_keys = torch.empty(n_isects, dtype=torch.int32, device=dev)      # 16- or 32-bit keys inside
    flatten_ids = torch.empty(n_isects, dtype=torch.int32, device=dev)
    offsets = torch.empty(C, tile_height, tile_width, dtype=torch.int32, device=dev)
    ws2 = _ws(lib.b2s_tile_bin_workspace_bytes(M, n_isects), dev)
    with _stage("tile_bin"):
        check(lib.b2s_tile_bin(M, N, ptr(cam), n_isects, ptr(sorted_ids), ptr(means2d), ptr(radii), C, tile_size, tile_width,
                               tile_height, ptr(tile_keys), ptr(flatten_ids), ptr(offsets), ptr(ws2), ws2.numel(),
                               L.tile_rows(tile_rows), st), "b2s_tile_bin")

    def make_isect_ids():
        ids = torch.empty(n_isects, dtype=torch.int64, device=dev)
        check(lib.b2s_isect_ids_from_bins(n_isects, key_bytes, ptr(tile_keys), ptr(flatten_ids), ptr(depths), ptr(ids),
                                          L.stream_ptr(dev)), "b2s_isect_ids_from_bins")
        return ids

    return flatten_ids, offsets, make_isect_ids


@torch.no_grad()
def isect_offset_encode(isect_ids: Tensor, n_cameras: int, tile_width: int, tile_height: int) -> Tensor:
    """gsplat ``isect_offset_encode``: int32 [C, tile_height, tile_width]."""
    dev = L.require_cuda(isect_ids)
    isect_ids = isect_ids.contiguous()
    offsets = torch.empty(n_cameras, tile_height, tile_width, dtype=torch.int32, device=dev)
    with _stage("isect_offsets"):
        check(L.lib().b2s_isect_offsets(isect_ids.numel(), ptr(isect_ids), n_cameras, tile_width, tile_height, ptr(offsets),
                                        L.stream_ptr(dev)), "b2s_isect_offsets")
    return offsets


# ======================================================================================
# two-level binning (what rasterization() runs; include/b2s.h "Two-level binning")
# ======================================================================================
COARSE = 8   # B2S_COARSE: tiles per coarse-bin side


def coarse_bins_supported(n_cameras: int, tile_width: int, tile_height: int, channels: int) -> bool:
    bw, bh = (tile_width + COARSE - 1) // COARSE, (tile_height + COARSE - 1) // COARSE
    return n_cameras * bw * bh <= 65536 and channels in (1, 2, 3, 4)


class CoarseBins:
    """Depth-ordered rows binned into 8x8-tile coarse bins; the compositing kernels expand per-tile lists from it."""
    __slots__ = ("recs", "ids", "offsets", "sorted_ids", "n_records", "n_isects", "n_visible")

    def __init__(self, recs, ids, offsets, sorted_ids, n_records, n_isects, n_visible=None):
        self.recs, self.ids, self.offsets, self.sorted_ids = recs, ids, offsets, sorted_ids
        self.n_records, self.n_isects, self.n_visible = n_records, n_isects, n_visible


@torch.no_grad()
def bin_coarse(means2d: Tensor, radii: Tensor, depths: Tensor, tile_size: int, tile_width: int, tile_height: int,
               n_cameras: int, camera_ids: Optional[Tensor] = None, depth_keys: Optional[Tensor] = None,
               row_ids: Optional[Tensor] = None, totals: Optional[Tensor] = None,
               tile_rows: Optional[Tuple[int, int]] = None, counts: Optional[Tuple[int, int]] = None) -> CoarseBins:
    """Depth sort of the rows + coarse binning (b2s_depth_order, b2s_coarse_bin).  ``depth_keys`` / ``row_ids`` /
    ``totals`` ([B2S_TOTAL_SLOTS,4] counters) come from stage 1 when it ran fused with the same rows."""
    dev = L.require_cuda(means2d, radii, depths)
    packed = camera_ids is not None
    M = radii.numel()
    N = 0 if packed else means2d.shape[1]
    C = int(n_cameras)
    means2d, depths = f32c(means2d), f32c(depths)
    radii = radii.contiguous()
    cam = camera_ids.contiguous().to(torch.int64) if packed else None
    lib, st = L.lib(), L.stream_ptr(dev)
    if depth_keys is None or row_ids is None:
        depth_keys = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        row_ids = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        with _stage("isect_depth_keys"):
            check(lib.b2s_isect_depth_keys(M, ptr(depths), ptr(radii), ptr(depth_keys), ptr(row_ids), st), "b2s_isect_depth_keys")
    if totals is not None and tile_rows is not None:
        raise B2SError("stage-1 totals count the whole grid: pass totals=None with a tile-row window")
    if totals is None and counts is None:
        tot = torch.zeros(1, dtype=torch.int64, device=dev)
        with _stage("coarse_count"):
            check(lib.b2s_coarse_count(M, ptr(means2d), ptr(radii), tile_size, tile_width, tile_height, ptr(tot),
                                       L.tile_rows(tile_rows), st), "b2s_coarse_count")
    # The one host sync of a forward (the counts size the record buffers).  The counts are complete as soon as stage 1 is:
    # they are copied to pinned host memory on a second stream NOW, the depth sort is enqueued behind them on the main
    # stream, and the host waits for the copy only -- so it sizes, allocates and enqueues the rest of the forward while the
    # sort runs.  (A plain `totals.tolist()` after the sort's launch is ordered behind the sort and leaves the GPU idle for
    # as long as the host then needs to enqueue the next kernels.)
    totals_host = totals_ready = None
    if totals is not None and counts is None and os.environ.get("B2S_ASYNC_COUNTS", "1") != "0":
        totals_host, totals_ready = _async_to_host(totals)
    sorted_ids = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
    # fused path: the sort's histogram kernel (a streaming pass over the rows in row order) also stores every row's packed tile
    # rectangle by row; the one-kernel expansion gathers one word per depth-ordered row (B2S_BINNING_FUSED=0: the unfused
    # reference kernels, same outputs bit for bit; they also serve tile grids beyond 255 x 255)
    fused = os.environ.get("B2S_BINNING_FUSED", "1") != "0" and tile_width <= 255 and tile_height <= 255
    if fused:
        rects = torch.empty(max(M, 1), dtype=torch.int32, device=dev)
        ws = _ws(lib.b2s_depth_order_rects_workspace_bytes(M), dev)
        with _stage("depth_order", launches=6):      # histogram, scan, 4 digit passes
            check(lib.b2s_depth_order_rects(M, ptr(depth_keys), ptr(row_ids), ptr(means2d), ptr(radii), tile_size, tile_width,
                                            tile_height, L.tile_rows(tile_rows), ptr(sorted_ids), ptr(rects), ptr(ws), ws.numel(),
                                            st), "b2s_depth_order_rects")
    else:
        ws = _ws(lib.b2s_depth_order_workspace_bytes(M), dev)
        with _stage("depth_order", launches=6):      # histogram, scan, 4 digit passes
            check(lib.b2s_depth_order(M, ptr(depth_keys), ptr(row_ids), ptr(sorted_ids), ptr(ws), ws.numel(), st), "b2s_depth_order")
    n_visible = None
    if counts is not None:             # (n_isects, n_records) already on the host (the multi-GPU stripe table)
        n_isects, n_records = int(counts[0]), int(counts[1])
    elif totals is not None:
        if totals_ready is not None:
            totals_ready.synchronize()
            slots = totals_host.tolist()           # 32 counter rows, added up on the host (no reduction kernel)
        else:
            slots = totals.tolist()                # B2S_ASYNC_COUNTS=0: ordered behind the depth sort (A/B)
        n_isects, n_visible, n_records = (int(sum(r[k] for r in slots)) for k in (0, 1, 2))
    else:
        n_isects, n_records = None, int(tot.item())
    if n_records >= 2 ** 31:
        raise B2SError(f"{n_records} coarse-bin records do not fit the int32 offsets of the API")
    bw, bh = (tile_width + COARSE - 1) // COARSE, (tile_height + COARSE - 1) // COARSE
    n_alloc = (n_records + 3) // 4 * 4 + 4
    recs = torch.empty(n_alloc, dtype=torch.int32, device=dev)
    ids = torch.empty(n_alloc, dtype=torch.int32, device=dev)
    offsets = torch.empty(C * bw * bh + 1, dtype=torch.int32, device=dev)
    two_passes = (C * bw * bh - 1).bit_length() > 8
    if fused:
        ws2 = _ws(lib.b2s_coarse_bin_rects_workspace_bytes(M, n_records), dev)
        with _stage("coarse_bin", launches=4 + two_passes):          # emit, scan, passes, offsets
            check(lib.b2s_coarse_bin_rects(M, N, ptr(cam), n_records, ptr(sorted_ids), ptr(rects), C, tile_width,
                                           tile_height, ptr(recs), ptr(ids), ptr(offsets), ptr(ws2), ws2.numel(), st),
                  "b2s_coarse_bin_rects")
        return CoarseBins(recs, ids, offsets, sorted_ids, n_records, n_isects, n_visible)
    ws2 = _ws(lib.b2s_coarse_bin_workspace_bytes(M, n_records), dev)
    with _stage("coarse_bin", launches=7 + two_passes):   # rects, scan, emit, histogram, scan, passes, offsets
        check(lib.b2s_coarse_bin(M, N, ptr(cam), n_records, ptr(sorted_ids), ptr(means2d), ptr(radii), C, tile_size, tile_width,
                                 tile_height, ptr(recs), ptr(ids), ptr(offsets), ptr(ws2), ws2.numel(), L.tile_rows(tile_rows),
                                 st), "b2s_coarse_bin")
    return CoarseBins(recs, ids, offsets, sorted_ids, n_records, n_isects, n_visible)


# ======================================================================================
# stage 3: compositing
# ======================================================================================
_CHANNELS = (1, 2, 3, 4, 8, 16, 32)


class _RasterizeToPixels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2d, conics, colors, opacities, backgrounds, width, height, tile_size, isect_offsets,
                flatten_ids, absgrad, bins=None, n_cameras=None, tile_rows=None):
        """Either the reference's lists (isect_offsets [C,th,tw], flatten_ids [I]) or ``bins`` (CoarseBins, with
        ``n_cameras``): then the kernels expand the per-tile lists from the coarse bins themselves.
        ``tile_rows=(lo, hi)``: composite only tile rows [lo, hi); the outputs hold just those pixel rows."""
        dev = L.require_cuda(means2d, conics, colors, opacities)
        means2d, conics, colors, opacities = f32c(means2d), f32c(conics), f32c(colors), f32c(opacities)
        backgrounds = f32c(backgrounds)
        if bins is not None:
            C, tw, th = int(n_cameras), (width + tile_size - 1) // tile_size, (height + tile_size - 1) // tile_size
            isect_offsets = bins.offsets            # (saved in the slots of the classic lists)
            flatten_ids = bins.ids
            recs = bins.recs
            n_list = bins.n_records
        else:
            C, th, tw = isect_offsets.shape
            recs = None
            isect_offsets = isect_offsets.contiguous()
            flatten_ids = flatten_ids.contiguous()
            n_list = flatten_ids.numel()
        ch = colors.shape[-1]
        row_lo, row_hi = (0, th) if tile_rows is None else (int(tile_rows[0]), int(tile_rows[1]))
        if not 0 <= row_lo <= row_hi <= th:
            raise B2SError(f"tile_rows {tile_rows} outside the {th}-row tile grid")
        out_h = max(0, min(height, row_hi * tile_size) - row_lo * tile_size)
        render = torch.empty(C, out_h, width, ch, dtype=torch.float32, device=dev)
        alphas = torch.empty(C, out_h, width, 1, dtype=torch.float32, device=dev)
        last_ids = torch.empty(C, out_h, width, dtype=torch.int32, device=dev)
        final_T = torch.empty(C, out_h, width, dtype=torch.float32, device=dev)
        a = L.CompositeFwdArgs(
            C=C, width=width, height=height, tile_size=tile_size, tile_w=tw, tile_h=th, channels=ch,
            n_isects=n_list, means2d=ptr(means2d), conics=ptr(conics), colors=ptr(colors),
            opacities=ptr(opacities), backgrounds=ptr(backgrounds),
            isect_offsets=None if bins is not None else ptr(isect_offsets),
            flatten_ids=None if bins is not None else ptr(flatten_ids),
            render_colors=ptr(render), render_alphas=ptr(alphas), last_ids=ptr(last_ids), final_T=ptr(final_T),
            bin_offsets=ptr(isect_offsets) if bins is not None else None, bin_recs=ptr(recs),
            bin_ids=ptr(flatten_ids) if bins is not None else None, tile_row_lo=row_lo, tile_row_hi=row_hi)
        ctx.slab_pre = None
        if _prefill_enabled():
            flush_prefills(dev)                 # stage 1's deferred fill starts now, beside the compositing kernel
            if any(ctx.needs_input_grad[:4]):
                # the slab of atomically accumulated gradients, zero-filled on the side stream under this forward
                ctx.slab_pre = prefill_zeros(opacities.numel() * (6 + ch + (2 if absgrad else 0)), dev)
        # an empty window (a rank whose stripe holds no tile row) launches nothing, forward or backward; (0, 0) would
        # read as "no window" in the C ABI, so it never gets there
        if row_hi > row_lo:
            with _stage("composite_fwd"):
                check(L.lib().b2s_composite_fwd(ctypes.byref(a), L.stream_ptr(dev)), "b2s_composite_fwd")
        ctx.save_for_backward(means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids, alphas,
                              last_ids, final_T, recs)
        ctx.cfg = (width, height, tile_size, absgrad, C, tw, th, n_list, row_lo, row_hi)
        return render, alphas

    @staticmethod
    def backward(ctx, v_render, v_alphas):
        (means2d, conics, colors, opacities, backgrounds, isect_offsets, flatten_ids, alphas, last_ids,
         final_T, recs) = ctx.saved_tensors
        width, height, tile_size, absgrad, C, tw, th, n_list, row_lo, row_hi = ctx.cfg
        dev = means2d.device
        use_bins = recs is not None
        ch = colors.shape[-1]
        v_render, v_alphas = f32c(v_render), f32c(v_alphas)
        # one zero-filled slab for every atomically accumulated gradient (one fill instead of four)
        M = opacities.numel()
        widths = [2, 3, ch, 1] + ([2] if absgrad else [])
        if M * max(widths) >= 2 ** 32:
            raise B2SError(f"{M} rows x {max(widths)} gradient columns exceed the 32-bit element offsets of b2s_composite_bwd")
        if ctx.slab_pre is not None:
            pre, ctx.slab_pre = ctx.slab_pre, None       # (a second backward of the same graph gets a fresh slab)
            slab = take_prefilled(pre, M * sum(widths), dev)
        else:
            slab = zeros_f32(M * sum(widths), dev)
        parts, o = [], 0
        for wd in widths:
            parts.append(slab[o:o + M * wd])
            o += M * wd
        v_means2d, v_conics = parts[0].view(means2d.shape), parts[1].view(conics.shape)
        v_colors, v_opac = parts[2].view(colors.shape), parts[3].view(opacities.shape)
        v_abs = parts[4].view(means2d.shape) if absgrad else None
        a = L.CompositeBwdArgs(
            C=C, width=width, height=height, tile_size=tile_size, tile_w=tw, tile_h=th, channels=ch,
            n_isects=n_list, means2d=ptr(means2d), conics=ptr(conics), colors=ptr(colors),
            opacities=ptr(opacities), backgrounds=ptr(backgrounds),
            isect_offsets=None if use_bins else ptr(isect_offsets), flatten_ids=None if use_bins else ptr(flatten_ids),
            render_alphas=ptr(alphas), last_ids=ptr(last_ids), final_T=ptr(final_T),
            v_render_colors=ptr(v_render), v_render_alphas=ptr(v_alphas), v_means2d=ptr(v_means2d),
            v_means2d_abs=ptr(v_abs), v_conics=ptr(v_conics), v_colors=ptr(v_colors), v_opacities=ptr(v_opac),
            bin_offsets=ptr(isect_offsets) if use_bins else None, bin_recs=ptr(recs),
            bin_ids=ptr(flatten_ids) if use_bins else None, tile_row_lo=row_lo, tile_row_hi=row_hi)
        if row_hi > row_lo:                         # empty window: the zero-filled slab is the gradient
            with _stage("composite_bwd"):
                check(L.lib().b2s_composite_bwd(ctypes.byref(a), L.stream_ptr(dev)), "b2s_composite_bwd")
        v_bg = None
        if backgrounds is not None and ctx.needs_input_grad[4]:
            v_bg = (v_render * final_T.abs()[..., None]).sum(dim=(1, 2))
        if absgrad:
            getattr(ctx, "absgrad_holder", means2d).absgrad = v_abs
        return v_means2d, v_conics, v_colors, v_opac, v_bg, None, None, None, None, None, None, None, None, None


def rasterize_to_pixels(means2d: Tensor, conics: Tensor, colors: Tensor, opacities: Tensor, image_width: int,
                        image_height: int, tile_size: int, isect_offsets: Tensor, flatten_ids: Tensor,
                        backgrounds: Optional[Tensor] = None, masks: Optional[Tensor] = None, packed: bool = False,
                        absgrad: bool = False, bins: Optional["CoarseBins"] = None,
                        n_cameras: Optional[int] = None, tile_rows: Optional[Tuple[int, int]] = None) -> Tuple[Tensor, Tensor]:
    """gsplat ``rasterize_to_pixels``: (render_colors [C,H,W,ch], render_alphas [C,H,W,1]).  Rows of the
    per-splat tensors are addressed through ``flatten_ids`` so dense ``[C,N,..]`` and packed ``[nnz,..]`` both work."""
    if masks is not None:
        raise NotImplementedError("tile masks are not implemented")
    if tile_size != 16:
        raise NotImplementedError("tile_size must be 16")
    C = int(n_cameras) if bins is not None else isect_offsets.shape[0]
    if packed:
        if means2d.dim() != 2:
            raise ValueError("packed=True expects [nnz,..] tensors")
    elif means2d.dim() != 3 or means2d.shape[0] != C:
        raise ValueError("dense layout expects [C,N,..] tensors")
    ch = colors.shape[-1]
    pad = 0
    if ch not in _CHANNELS:
        if ch > 32:
            raise NotImplementedError("more than 32 channels: chunk them (rasterization() does)")
        target = next(c for c in _CHANNELS if c >= ch)
        pad = target - ch
        colors = torch.cat([colors, colors.new_zeros(colors.shape[:-1] + (pad,))], dim=-1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, backgrounds.new_zeros(backgrounds.shape[:-1] + (pad,))], dim=-1)
    holder = means2d
    fn = _RasterizeToPixels
    # absgrad: gsplat attaches `.absgrad` to the means2d tensor the caller passed in
    render, alphas = _apply_with_holder(fn, holder, means2d, conics, colors, opacities, backgrounds, int(image_width),
                                        int(image_height), int(tile_size), isect_offsets, flatten_ids, bool(absgrad), bins, C,
                                        tile_rows)
    if pad:
        render = render[..., :ch]
    return render, alphas


def _apply_with_holder(fn, holder, *args):
    """Runs ``fn.apply`` and lets its backward find the tensor that should receive ``.absgrad``."""
    out = fn.apply(*args)
    node = out[0].grad_fn
    if node is not None:
        node.absgrad_holder = holder
    return out
