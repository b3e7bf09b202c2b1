"""GPU: the fused two-level binning (b2s_depth_order_rects + b2s_coarse_bin_rects: the depth sort's last pass stores the
tile rectangles in depth order, one kernel expands them into coarse-bin records and counts the digits of the bin sort)
against the unfused kernels (b2s_depth_order + b2s_coarse_bin) and against a record list built here, independently, from the
oracle's tile rectangles.  Integer work: everything bit-exact."""
import pytest
import torch

from oracle import torch_oracle as O
from tests.util import scene_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2s():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pipeline_pointcloud_b200 as pkg
    return pkg


def project(b2s, N, C, W, H, seed, stress=False, scale=1.0):
    a = scene_args(N, 0, C, W, H, seed=seed, stress=stress)
    g = {k: v.cuda() for k, v in a.items()}
    radii, means2d, depths, _, _ = b2s.fully_fused_projection(g["means"], None, g["quats"], g["scales"] * scale, g["viewmats"],
                                                              g["Ks"], W, H)
    return radii, means2d, depths


def both(monkeypatch, radii, means2d, depths, tw, th, C, **kw):
    from pipeline_pointcloud_b200 import ops
    out = []
    for mode in ("0", "1"):
        monkeypatch.setenv("B2S_BINNING_FUSED", mode)
        b = ops.bin_coarse(means2d, radii, depths, 16, tw, th, C, **kw)
        n = b.n_records
        out.append((b.recs[:n].clone(), b.ids[:n].clone(), b.offsets.clone(), b.sorted_ids.clone(), n))
    return out


def reference_records(radii, means2d, depths, tw, th, C, rows=None):
    """(rec, id) lists from first principles on the CPU: rows in ascending (depth bits, flat index) order -- culled rows last --
    each emitting one record per coarse bin its tile rectangle overlaps (bins row-major inside the rectangle), then a stable
    sort by coarse key."""
    r, m, d = radii.cpu().reshape(-1), means2d.cpu().reshape(-1, 2), depths.cpu().reshape(-1)
    M, N = r.numel(), r.numel() // C
    lo, hi = (0, th) if rows is None else rows
    tmin, tmax = O.tile_rects(m.reshape(1, M, 2), r.reshape(1, M), 16, tw, th)        # [1,M,2] each, (x, y)
    x0, y0, x1, y1 = tmin[0, :, 0], tmin[0, :, 1].clamp(lo, hi), tmax[0, :, 0], tmax[0, :, 1].clamp(lo, hi)
    keys = d.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = torch.where(r > 0, keys, torch.full_like(keys, 0xFFFFFFFF))
    order = torch.sort(keys, stable=True).indices
    bw, bh = (tw + 7) // 8, (th + 7) // 8
    recs, ids = [], []
    for i in order.tolist():
        if r[i] <= 0 or x1[i] <= x0[i] or y1[i] <= y0[i]:
            continue
        cam = i // N
        for by in range(int(y0[i]) // 8, (int(y1[i]) + 7) // 8):
            for bx in range(int(x0[i]) // 8, (int(x1[i]) + 7) // 8):
                mx = sum(1 << t for t in range(8) if int(x0[i]) <= bx * 8 + t < int(x1[i]))
                my = sum(1 << t for t in range(8) if int(y0[i]) <= by * 8 + t < int(y1[i]))
                recs.append((my << 24) | (mx << 16) | ((cam * bh + by) * bw + bx))
                ids.append(i)
    recs, ids = torch.tensor(recs, dtype=torch.int64), torch.tensor(ids, dtype=torch.int64)
    perm = torch.sort(recs & 0xFFFF, stable=True).indices
    recs, ids = recs[perm], ids[perm]
    offsets = torch.searchsorted((recs & 0xFFFF).contiguous(), torch.arange(C * bw * bh + 1))
    return recs, ids, offsets, order


@pytest.mark.parametrize("N,C,W,H,rows", [(3000, 1, 300, 200, None), (2500, 3, 203, 150, None), (3000, 2, 203, 150, (4, 9)),
                                          (2000, 1, 640, 400, (0, 0))])
def test_fused_binning_matches_first_principles(b2s, monkeypatch, N, C, W, H, rows):
    tw, th = (W + 15) // 16, (H + 15) // 16
    radii, means2d, depths = project(b2s, N, C, W, H, seed=5, stress=True)
    kw = {} if rows is None else {"tile_rows": rows}
    (ru, iu, ou, su, nu), (rf, i_f, of, sf, nf) = both(monkeypatch, radii, means2d, depths, tw, th, C, **kw)
    ref_r, ref_i, ref_o, ref_order = reference_records(radii, means2d, depths, tw, th, C, rows)
    assert nu == nf == ref_r.numel()
    for recs, ids, offs, srt in ((ru, iu, ou, su), (rf, i_f, of, sf)):
        assert torch.equal(recs.cpu().to(torch.int64) & 0xFFFFFFFF, ref_r)
        assert torch.equal(ids.cpu().to(torch.int64), ref_i)
        assert torch.equal(offs.cpu().to(torch.int64), ref_o)
        assert torch.equal(srt.cpu().to(torch.int64), ref_order)


@pytest.mark.parametrize("N,C,W,H,scale", [(400_000, 1, 1297, 840, 1.0),      # config-3 geometry: 77 bins, one digit pass
                                           (150_000, 8, 640, 400, 1.0),       # 8 cameras: 120 bins
                                           (300_000, 1, 3840, 2160, 1.0),     # config-5 geometry: 510 bins, two digit passes
                                           (60_000, 2, 1297, 840, 8.0),       # large splats: tens of records per row
                                           (1, 1, 64, 64, 1.0)])
def test_fused_binning_equals_unfused(b2s, monkeypatch, N, C, W, H, scale):
    tw, th = (W + 15) // 16, (H + 15) // 16
    radii, means2d, depths = project(b2s, N, C, W, H, seed=9, scale=scale)
    (ru, iu, ou, su, nu), (rf, i_f, of, sf, nf) = both(monkeypatch, radii, means2d, depths, tw, th, C)
    assert nu == nf
    assert torch.equal(su, sf) and torch.equal(ou, of) and torch.equal(ru, rf) and torch.equal(iu, i_f)
    assert int(of[-1]) == nf and bool((of[1:] >= of[:-1]).all())
    # a tile-row window (a multi-GPU stripe), including one that is empty
    for rows in ((th // 3, 2 * th // 3), (th // 2, th // 2)):
        (ru, iu, ou, su, nu), (rf, i_f, of, sf, nf) = both(monkeypatch, radii, means2d, depths, tw, th, C, tile_rows=rows)
        assert nu == nf and torch.equal(su, sf) and torch.equal(ou, of) and torch.equal(ru, rf) and torch.equal(iu, i_f)


def test_fused_binning_renders_bit_identically(b2s, monkeypatch):
    from tests.util import call
    N, W, H = 200_000, 640, 400
    a = scene_args(N, 3, 2, W, H, seed=3)
    g = {k: v.cuda() for k, v in a.items()}
    out = {}
    with torch.no_grad():
        for mode in ("0", "1"):
            monkeypatch.setenv("B2S_BINNING_FUSED", mode)
            out[mode] = call(b2s.rasterization, g, W, H, sh_degree=3, packed=False, rasterize_mode="antialiased")[:2]
    assert torch.equal(out["0"][0], out["1"][0]) and torch.equal(out["0"][1], out["1"][1])
