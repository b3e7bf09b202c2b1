"""ctypes binding of the C-ABI library ``libb2s.so`` (``include/b2s.h``).

There is no fallback: if the library is missing or a call fails this raises.  The product never
imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2S_LIB") or os.path.join(_HERE, "libb2s.so")     # B2S_LIB: an A/B build of the same ABI

_lib = None


class B2SError(RuntimeError):
    pass


class TileRows(Structure):
    """b2s_tile_rows: tile-row window [row_lo, row_hi) of the tile grid (None / NULL = the whole grid)."""
    _fields_ = [("row_lo", c_int32), ("row_hi", c_int32)]


def tile_rows(rows):
    """ctypes argument for an optional (row_lo, row_hi) window."""
    return None if rows is None else ctypes.byref(TileRows(int(rows[0]), int(rows[1])))


class Stripes(Structure):
    """b2s_stripes: tile-row bounds of the ranks' stripes."""
    _fields_ = [("world", c_int32), ("cams_per_rank", c_int32), ("bounds", c_int32 * 17)]


class ProjectFwdArgs(Structure):
    _fields_ = [
        ("N", c_int64), ("C", c_int32),
        ("means", c_void_p), ("quats", c_void_p), ("scales", c_void_p), ("opacities", c_void_p),
        ("viewmats", c_void_p), ("Ks", c_void_p),
        ("width", c_int32), ("height", c_int32),
        ("eps2d", c_float), ("near_plane", c_float), ("far_plane", c_float), ("radius_clip", c_float),
        ("antialiased", c_int32),
        ("sh_coeffs", c_void_p), ("campos", c_void_p), ("sh_degree", c_int32), ("sh_stride_k", c_int32),
        ("tile_size", c_int32), ("tile_w", c_int32), ("tile_h", c_int32),
        ("radii", c_void_p), ("means2d", c_void_p), ("depths", c_void_p), ("conics", c_void_p),
        ("compensations", c_void_p), ("opacities_out", c_void_p), ("colors", c_void_p),
        ("tiles_per_gauss", c_void_p), ("depth_keys", c_void_p), ("row_ids", c_void_p),
        ("total_isects", c_void_p),
    ]


class RouteArgs(Structure):
    """b2s_route_args: stage 1 fused with the multi-GPU exchange (csrc/route.cu)."""
    _fields_ = [
        ("proj", ProjectFwdArgs),
        ("world", c_int32), ("rank", c_int32), ("cams_per_rank", c_int32), ("region_cams", c_int32),
        ("channels", c_int32), ("color_mode", c_int32), ("n_color_in", c_int32), ("depth_channel", c_int32),
        ("overflow", c_int32), ("bounds", c_int32 * 17),
        ("cap_rows", c_int64), ("colors_in", c_void_p),
        ("peer_rec", c_void_p * 16), ("peer_cnt", c_void_p * 16),
    ]


class ProjectBwdArgs(Structure):
    _fields_ = [
        ("N", c_int64), ("C", c_int32),
        ("means", c_void_p), ("quats", c_void_p), ("scales", c_void_p), ("opacities", c_void_p),
        ("viewmats", c_void_p), ("Ks", c_void_p),
        ("width", c_int32), ("height", c_int32),
        ("eps2d", c_float), ("near_plane", c_float), ("far_plane", c_float), ("radius_clip", c_float),
        ("antialiased", c_int32),
        ("sh_coeffs", c_void_p), ("campos", c_void_p), ("sh_degree", c_int32), ("sh_stride_k", c_int32),
        ("radii", c_void_p),
        ("v_means2d", c_void_p), ("v_depths", c_void_p), ("v_conics", c_void_p),
        ("v_compensations", c_void_p), ("v_opacities_out", c_void_p), ("v_colors", c_void_p),
        ("accumulate", c_int32),
        ("v_means", c_void_p), ("v_quats", c_void_p), ("v_scales", c_void_p), ("v_opacities", c_void_p),
        ("v_sh_coeffs", c_void_p), ("stages", c_int32),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
        ("v_sh_prezeroed", c_int32),
    ]


class CompositeFwdArgs(Structure):
    _fields_ = [
        ("C", c_int32), ("width", c_int32), ("height", c_int32), ("tile_size", c_int32),
        ("tile_w", c_int32), ("tile_h", c_int32), ("channels", c_int32),
        ("n_isects", c_int64),
        ("means2d", c_void_p), ("conics", c_void_p), ("colors", c_void_p), ("opacities", c_void_p),
        ("backgrounds", c_void_p), ("isect_offsets", c_void_p), ("flatten_ids", c_void_p),
        ("render_colors", c_void_p), ("render_alphas", c_void_p), ("last_ids", c_void_p), ("final_T", c_void_p),
        ("bin_offsets", c_void_p), ("bin_recs", c_void_p), ("bin_ids", c_void_p),
        ("tile_row_lo", c_int32), ("tile_row_hi", c_int32),
    ]


class CompositeBwdArgs(Structure):
    _fields_ = [
        ("C", c_int32), ("width", c_int32), ("height", c_int32), ("tile_size", c_int32),
        ("tile_w", c_int32), ("tile_h", c_int32), ("channels", c_int32),
        ("n_isects", c_int64),
        ("means2d", c_void_p), ("conics", c_void_p), ("colors", c_void_p), ("opacities", c_void_p),
        ("backgrounds", c_void_p), ("isect_offsets", c_void_p), ("flatten_ids", c_void_p),
        ("render_alphas", c_void_p), ("last_ids", c_void_p), ("final_T", c_void_p),
        ("v_render_colors", c_void_p), ("v_render_alphas", c_void_p),
        ("v_means2d", c_void_p), ("v_means2d_abs", c_void_p), ("v_conics", c_void_p), ("v_colors", c_void_p),
        ("v_opacities", c_void_p), ("bin_offsets", c_void_p), ("bin_recs", c_void_p), ("bin_ids", c_void_p),
        ("tile_row_lo", c_int32), ("tile_row_hi", c_int32),
    ]


# every symbol include/b2s.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b2s_version": (c_int32, []),
    "b2s_last_error": (c_char_p, []),
    "b2s_project_fwd": (c_int32, [POINTER(ProjectFwdArgs), c_void_p]),
    "b2s_zero_fill": (c_int32, [c_void_p, c_int64, c_void_p]),
    "b2s_project_bwd_workspace_bytes": (c_int64, [c_int64]),
    "b2s_project_bwd": (c_int32, [POINTER(ProjectBwdArgs), c_void_p]),
    "b2s_viewmat_bwd": (c_int32, [POINTER(ProjectBwdArgs), c_void_p, c_void_p]),
    "b2s_sh_fwd": (c_int32, [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_sh_bwd": (c_int32, [c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_quat_scale_to_covar_preci": (c_int32, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_isect_count": (c_int32, [c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, POINTER(TileRows), c_void_p]),
    "b2s_isect_scan_workspace_bytes": (c_int64, [c_int64]),
    "b2s_isect_scan": (c_int32, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_select_visible": (c_int32, [c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_isect_fill": (c_int32, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                 c_int32, c_void_p, c_void_p, POINTER(TileRows), c_void_p]),
    "b2s_isect_sort_workspace_bytes": (c_int64, [c_int64]),
    "b2s_isect_sort": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_isect_offsets": (c_int32, [c_int64, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "b2s_isect_depth_keys": (c_int32, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_depth_order_workspace_bytes": (c_int64, [c_int64]),
    "b2s_depth_order": (c_int32, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_tile_key_bytes": (c_int32, [c_int32, c_int32, c_int32]),
    "b2s_tile_bin_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "b2s_tile_bin": (c_int32, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                               c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(TileRows), c_void_p]),
    "b2s_isect_ids_from_bins": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_coarse_count": (c_int32, [c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, POINTER(TileRows), c_void_p]),
    "b2s_depth_order_rects_workspace_bytes": (c_int64, [c_int64]),
    "b2s_depth_order_rects": (c_int32, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                        POINTER(TileRows), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_coarse_bin_rects_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "b2s_coarse_bin_rects": (c_int32, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32,
                                       c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_coarse_bin_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "b2s_coarse_bin": (c_int32, [c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                 c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, POINTER(TileRows), c_void_p]),
    "b2s_pack_rows": (c_int32, [c_int64, c_int64, c_void_p, c_int32] + [c_void_p] * 17),
    "b2s_unpack_rows": (c_int32, [c_int64, c_void_p, c_int32] + [c_void_p] * 11),
    "b2s_pack_payload": (c_int32, [c_int64, c_int64, c_void_p, c_int32, c_int32, c_int64] + [c_void_p] * 9),
    "b2s_split_payload": (c_int32, [c_int64, c_int32] + [c_void_p] * 11),
    "b2s_join_grads": (c_int32, [c_int64, c_int32] + [c_void_p] * 7),
    "b2s_scatter_grads": (c_int32, [c_int64, c_void_p, c_int32] + [c_void_p] * 6 + [c_int32, c_void_p]),
    "b2s_stripe_ctas": (c_int64, [c_int64]),
    "b2s_stripe_count": (c_int32, [c_int64, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, POINTER(Stripes), c_void_p,
                                   c_void_p, c_void_p, c_int64, c_void_p]),
    "b2s_stripe_scan": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "b2s_p2p_put": (c_int32, [c_int32, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "b2s_stripe_pack": (c_int32, [c_int64, c_int64, c_int32, c_int64] + [c_void_p] * 8 + [c_int32, c_int32, c_void_p, c_void_p,
                                                                                           c_void_p]),
    "b2s_stripe_unpack": (c_int32, [c_int64, c_int32, c_void_p, c_int32, c_int32, c_int32, POINTER(TileRows)] +
                          [c_void_p] * 11 + [c_void_p]),
    "b2s_stripe_return": (c_int32, [c_int64, c_int32] + [c_void_p] * 7 + [c_int64, c_int32] + [c_void_p] * 6 + [c_void_p]),
    "b2s_stripe_scatter_add": (c_int32, [c_int32, c_int32] + [c_void_p] * 9 + [c_void_p]),
    "b2s_grad_compact": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_void_p]),
    "b2s_grad_scatter_add": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_route_rec_floats": (c_int64, [c_int32, c_int32, c_int64]),
    "b2s_route_cnt_ints": (c_int64, [c_int32, c_int32, c_int64]),
    "b2s_route_project": (c_int32, [POINTER(RouteArgs), c_void_p]),
    "b2s_route_scan": (c_int32, [c_int32, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_route_gather": (c_int32, [c_int32, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                   c_int32, c_int32, c_int32, POINTER(TileRows)] + [c_void_p] * 11 + [c_void_p]),
    "b2s_l1_ssim_ctas": (c_int64, [c_int32, c_int32, c_int32]),
    "b2s_l1_ssim_fwd": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2s_l1_ssim_bwd": (c_int32, [c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p,
                                  c_void_p]),
    "b2s_adam_step": (c_int32, [c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                c_float, c_int64, c_void_p]),
    "b2s_spz_pack": (c_int32, [c_int64, c_int32] + [c_void_p] * 6 + [c_int32] + [c_void_p] * 6 + [c_void_p]),
    "b2s_spz_unpack": (c_int32, [c_int64, c_int32, c_int32] + [c_void_p] * 7 + [c_void_p]),
    "b2s_convert_spz_to_ply": (c_int32, [ctypes.c_char_p, ctypes.c_char_p]),
    "b2s_convert_ply_to_spz": (c_int32, [ctypes.c_char_p, ctypes.c_char_p]),
    "b2s_composite_fwd": (c_int32, [POINTER(CompositeFwdArgs), c_void_p]),
    "b2s_composite_bwd": (c_int32, [POINTER(CompositeBwdArgs), c_void_p]),
}


def lib() -> ctypes.CDLL:
    """Loads libb2s.so once.  Raises if it is not built -- there is no other code path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2SError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C pipeline_pointcloud_b200/csrc`).  There is no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b2s_last_error().decode("utf-8", "replace")
        raise B2SError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int:
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise B2SError("b2s runs on CUDA tensors only (there is no CPU path); got a tensor on " + str(t.device))
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise B2SError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise B2SError("no CUDA tensor given")
    return dev


def f32c(t):
    """Contiguous fp32 view/copy (None passes through)."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
