"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): tile-intersection counts and sorted key/index lists bit-exact; rendered
RGB/alpha within 1e-4 absolute; per-Gaussian gradients within 1e-3 relative; fp32.
Pixels whose compositing decision sits on a knife edge (alpha == 1/255 or T == 1e-4 to within a few ulp, where
`ex2.approx` vs `exp` legitimately flips a discrete decision) are reported by the oracle and excluded; their
fraction is asserted to stay small.
"""
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from tests.util import LEAVES, call, grad_report, scene_args

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
RENDER_ATOL = 1e-4


@pytest.fixture(scope="module")
def b2s():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pipeline_pointcloud_b200 as pkg
    return pkg


def bits(t):
    return t.detach().cpu().contiguous().view(torch.int32)


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,W,H", [(1, 200, 120), (3, 97, 61)])
def test_projection_bit_exact(b2s, C, W, H):
    a = scene_args(6000, 3, C, W, H, seed=1, stress=True)
    g = {k: v.cuda() for k, v in a.items()}
    radii, means2d, depths, conics, comps = b2s.fully_fused_projection(
        g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H, calc_compensations=True)
    o = O.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], W, H, calc_compensations=True)
    assert (o[0] > 0).sum() > 500 and (o[0] == 0).sum() > 50
    assert torch.equal(radii.cpu(), o[0])
    for name, x, y in zip(("means2d", "depths", "conics", "compensations"), (means2d, depths, conics, comps), o[1:]):
        assert torch.equal(bits(x), bits(y)), name


def test_projection_packed_layout(b2s):
    W, H = 120, 80
    a = scene_args(3000, 0, 2, W, H, seed=2, stress=True)
    g = {k: v.cuda() for k, v in a.items()}
    cam, gid, radii, means2d, depths, conics, comps = b2s.fully_fused_projection(
        g["means"], None, g["quats"], g["scales"], g["viewmats"], g["Ks"], W, H, packed=True)
    o = O.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], W, H)
    sel = (o[0].reshape(-1) > 0).nonzero()[:, 0]
    assert cam.dtype == torch.int64 and gid.dtype == torch.int64
    assert torch.equal(cam.cpu(), sel // 3000) and torch.equal(gid.cpu(), sel % 3000)
    assert torch.equal(bits(means2d), bits(o[1].reshape(-1, 2)[sel]))


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_colours(b2s, deg):
    torch.manual_seed(deg)
    dirs = torch.randn(5000, 3) * 2
    coeffs = torch.randn(5000, 16, 3)
    masks = torch.rand(5000) > 0.2
    ref = O.spherical_harmonics(deg, dirs.double(), coeffs.double(), masks).float()
    dg, cg = dirs.cuda().requires_grad_(), coeffs.cuda().requires_grad_()
    got = b2s.spherical_harmonics(deg, dg, cg, masks.cuda())
    assert (got.cpu() - ref).abs().max() < 2e-5
    v = torch.randn(5000, 3)
    got.backward(v.cuda())
    d64, c64 = dirs.double().requires_grad_(), coeffs.double().requires_grad_()
    (O.spherical_harmonics(deg, d64, c64, masks) * v).sum().backward()
    assert (cg.grad.cpu() - c64.grad.float()).abs().max() < 2e-5
    if deg > 0:
        assert grad_report(dg.grad, d64.grad)[0] < 1e-4


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,W,H,packed", [(1, 256, 256, False), (3, 203, 117, False), (2, 160, 96, True),
                                          (8, 1297, 840, False)])   # last: camera x tile ids need 17 bits (32-bit tile keys)
def test_binning_bit_exact(b2s, C, W, H, packed):
    N = 20000
    a = scene_args(N, 0, C, W, H, seed=3, stress=True)
    radii, means2d, depths, conics, _ = O.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    kw = {}
    if packed:
        sel = (radii.reshape(-1) > 0).nonzero()[:, 0]
        radii, means2d, depths = radii.reshape(-1)[sel], means2d.reshape(-1, 2)[sel], depths.reshape(-1)[sel]
        kw = dict(packed=True, n_cameras=C, camera_ids=sel // N, gaussian_ids=sel % N)
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, tw, th, **kw)
    off = O.isect_offset_encode(ids, C, tw, th)
    kwg = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
    assert ids.numel() > 10000
    for method in ("radix64", "depth_first"):     # the reference's algorithm and the low-traffic training path
        gtpg, gids, gflat = b2s.isect_tiles(means2d.cuda(), radii.cuda(), depths.cuda(), 16, tw, th, method=method, **kwg)
        goff = b2s.isect_offset_encode(gids, C, tw, th)
        assert torch.equal(gtpg.cpu(), tpg), method
        assert torch.equal(gids.cpu(), ids), method          # sorted (camera, tile, depth) keys, bit for bit
        assert torch.equal(gflat.cpu(), flat), method        # and the permuted indices (stable ties)
        assert torch.equal(goff.cpu(), off), method
    from pipeline_pointcloud_b200 import ops
    f2, off2, mk = ops.bin_depth_first(means2d.cuda(), radii.cuda(), depths.cuda(), gtpg, 16, tw, th, C,
                                       camera_ids=kwg.get("camera_ids"))
    assert torch.equal(f2.cpu(), flat) and torch.equal(off2.cpu(), off) and torch.equal(mk().cpu(), ids)
    # unsorted variant
    _, uids, uflat = b2s.isect_tiles(means2d.cuda(), radii.cuda(), depths.cuda(), 16, tw, th, sort=False, **kwg)
    _, oids, oflat = O.isect_tiles(means2d, radii, depths, 16, tw, th, sort=False, **kw)
    assert torch.equal(uids.cpu(), oids) and torch.equal(uflat.cpu(), oflat)


def test_binning_empty(b2s):
    means2d = torch.zeros(2, 10, 2).cuda()
    radii = torch.zeros(2, 10, dtype=torch.int32).cuda()
    depths = torch.ones(2, 10).cuda()
    for method in ("radix64", "depth_first"):
        tpg, ids, flat = b2s.isect_tiles(means2d, radii, depths, 16, 4, 3, method=method)
        assert ids.numel() == 0 and flat.numel() == 0 and tpg.sum().item() == 0
        off = b2s.isect_offset_encode(ids, 2, 4, 3)
        assert off.shape == (2, 3, 4) and (off == 0).all()


def test_binning_depth_ties(b2s):
    """Equal (tile, depth) keys must keep ascending flat index (stable), in both binning methods."""
    N = 5000
    g = torch.Generator().manual_seed(0)
    means2d = (torch.rand(1, N, 2, generator=g) * torch.tensor([128.0, 96.0]))
    radii = torch.randint(0, 14, (1, N), generator=g, dtype=torch.int32)
    depths = torch.randint(1, 6, (1, N), generator=g).float()          # only 5 distinct depths => many ties
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, 8, 6)
    for method in ("radix64", "depth_first"):
        _, gids, gflat = b2s.isect_tiles(means2d.cuda(), radii.cuda(), depths.cuda(), 16, 8, 6, method=method)
        assert torch.equal(gids.cpu(), ids) and torch.equal(gflat.cpu(), flat), method


# ----------------------------------------------------------------------------------------------------------
def _compare_render(got_r, got_a, ref_r, ref_a, fragile, max_fragile=0.05):
    ok = ~fragile
    assert fragile.float().mean().item() < max_fragile
    dr = (got_r.cpu() - ref_r).abs()
    da = (got_a.cpu() - ref_a).abs()
    assert dr[ok].max().item() <= RENDER_ATOL, dr[ok].max().item()
    assert da[ok].max().item() <= RENDER_ATOL, da[ok].max().item()
    # knife-edge pixels may flip one 1/255 contribution, never more
    assert dr.max().item() < 2e-2 and da.max().item() < 2e-2


@pytest.mark.parametrize("ch,bg", [(3, False), (3, True), (1, False), (4, True), (5, False), (16, False)])
def test_rasterize_to_pixels_forward(b2s, ch, bg):
    W, H, C, N = 150, 90, 2, 4000
    a = scene_args(N, 0, C, W, H, seed=4)
    radii, means2d, depths, conics, _ = O.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, tw, th)
    off = O.isect_offset_encode(ids, C, tw, th)
    g = torch.Generator().manual_seed(ch)
    colors = torch.rand(C, N, ch, generator=g)
    opac = a["opacities"][None].expand(C, N).contiguous()
    bgs = torch.rand(C, ch, generator=g) if bg else None
    ref_r, ref_a, fragile = O.rasterize_to_pixels(means2d, conics, colors, opac, W, H, 16, off, flat, bgs, return_fragile=True)
    got_r, got_a = b2s.rasterize_to_pixels(means2d.cuda(), conics.cuda(), colors.cuda(), opac.cuda(), W, H, 16, off.cuda(),
                                           flat.cuda(), backgrounds=bgs.cuda() if bg else None)
    assert got_r.shape == (C, H, W, ch) and got_a.shape == (C, H, W, 1)
    _compare_render(got_r, got_a, ref_r, ref_a, fragile)


@pytest.mark.parametrize("cfg", [
    dict(N=3000, deg=3, C=2, W=96, H=64, packed=False, mode="antialiased"),                       # two-level binning (default)
    dict(N=3000, deg=3, C=2, W=96, H=64, packed=False, mode="antialiased", binning="tiles"),      # reference-shaped lists
    dict(N=3000, deg=3, C=1, W=111, H=70, packed=True, mode="classic"),
    dict(N=2500, deg=1, C=3, W=80, H=48, packed=True, mode="antialiased", render_mode="RGB+ED", bg=True),
    dict(N=2500, deg=2, C=2, W=80, H=48, packed=False, mode="antialiased", render_mode="RGB+ED", bg=True),
    dict(N=2000, deg=None, C=2, W=64, H=64, packed=False, mode="classic", render_mode="RGB+D"),
    dict(N=2000, deg=None, C=2, W=64, H=64, packed=False, mode="classic", render_mode="RGB+D", binning="tiles"),
    dict(N=3000, deg=3, C=1, W=111, H=70, packed=True, mode="classic", binning="tiles"),
    dict(N=20000, deg=3, C=1, W=64, H=48, packed=False, mode="classic", dense=True),              # deep overdraw: tiles saturate early
    dict(N=6000, deg=1, C=2, W=300, H=200, packed=False, mode="classic", big=True),               # several coarse bins per camera
])
def test_rasterization_forward_backward(b2s, cfg, monkeypatch):
    """The whole drop-in call, forward and backward, against the oracle's autograd."""
    monkeypatch.setenv("B2S_BINNING", cfg.get("binning", "coarse"))
    N, deg, C, W, H = cfg["N"], cfg["deg"], cfg["C"], cfg["W"], cfg["H"]
    a = scene_args(N, deg if deg is not None else 0, C, W, H, seed=6, requires_grad=True, stress=not cfg.get("dense"))
    if cfg.get("dense"):     # opaque, large splats: most tiles saturate inside the first chunk
        with torch.no_grad():
            a["scales"].mul_(3.0)
            a["opacities"].copy_(a["opacities"] * 0.3 + 0.69)
    if deg is None:
        a["colors"] = torch.rand(N, 3, generator=torch.Generator().manual_seed(9)).requires_grad_()
    bgs = torch.rand(C, 3, generator=torch.Generator().manual_seed(10)) if cfg.get("bg") else None
    kw = dict(sh_degree=deg, packed=cfg["packed"], rasterize_mode=cfg["mode"], render_mode=cfg.get("render_mode", "RGB"))
    ref_r, ref_a, ref_m = call(O.rasterization, a, W, H, backgrounds=bgs, return_fragile=True, **kw)
    g = {k: (v.detach().cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}
    got_r, got_a, got_m = call(b2s.rasterization, g, W, H, backgrounds=bgs.cuda() if bgs is not None else None, **kw)

    # binning: bit-exact end to end (projection is canonical-order, so the keys agree too)
    assert torch.equal(got_m["radii"].cpu(), ref_m["radii"])
    assert torch.equal(got_m["tiles_per_gauss"].cpu(), ref_m["tiles_per_gauss"])
    assert torch.equal(got_m["isect_ids"].cpu(), ref_m["isect_ids"])
    assert torch.equal(got_m["flatten_ids"].cpu(), ref_m["flatten_ids"])
    assert torch.equal(got_m["isect_offsets"].cpu(), ref_m["isect_offsets"])
    if cfg["packed"]:
        assert torch.equal(got_m["camera_ids"].cpu(), ref_m["camera_ids"])
        assert torch.equal(got_m["gaussian_ids"].cpu(), ref_m["gaussian_ids"])
    assert torch.equal(bits(got_m["means2d"]), bits(ref_m["means2d"]))
    assert (got_m["opacities"].cpu() - ref_m["opacities"]).abs().max() < 1e-6

    fragile = ref_m["fragile"]
    if cfg.get("render_mode", "RGB").endswith("ED"):
        # expected depth divides by alpha: compare the colour channels at 1e-4 and depth relatively
        _compare_render(got_r[..., :3], got_a, ref_r[..., :3], ref_a, fragile)
        ok = ~fragile & (ref_a[..., 0] > 0.05)
        dd = (got_r.cpu()[..., 3] - ref_r[..., 3]).abs() / ref_r[..., 3].abs().clamp_min(1.0)
        assert dd[ok].max().item() < 1e-3
    elif cfg.get("render_mode") == "RGB+D":
        _compare_render(got_r[..., :3], got_a, ref_r[..., :3], ref_a, fragile)
        dd = (got_r.cpu()[..., 3] - ref_r[..., 3]).abs() / ref_r[..., 3].abs().clamp_min(1.0)
        assert dd[~fragile].max().item() < 1e-4
    else:
        _compare_render(got_r, got_a, ref_r, ref_a, fragile)

    # backward: random cotangent, zero on knife-edge pixels so both sides differentiate the same function
    gen = torch.Generator().manual_seed(12)
    keep = (~fragile)[..., None].float()
    v_r = torch.rand(ref_r.shape, generator=gen) * keep
    v_a = torch.rand(ref_a.shape, generator=gen) * keep
    ((ref_r * v_r).sum() + (ref_a * v_a).sum()).backward()
    got_m["means2d"].retain_grad()
    ((got_r * v_r.cuda()).sum() + (got_a * v_a.cuda()).sum()).backward()
    for k in LEAVES:
        rel, bad = grad_report(g[k].grad, a[k].grad)
        assert rel < 1e-3, (k, rel)        # north-star bar: 1e-3 relative
        assert bad < 5e-3, (k, bad)
    assert got_m["means2d"].grad is not None and got_m["means2d"].grad.abs().sum() > 0


def test_absgrad_and_means2d_grad(b2s):
    W, H = 64, 64
    a = scene_args(1500, 0, 1, W, H, seed=8)
    g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}
    r, al, m = call(b2s.rasterization, g, W, H, sh_degree=0, packed=False, absgrad=True)
    m["means2d"].retain_grad()
    r.sum().backward()
    assert m["means2d"].grad.shape == m["means2d"].shape
    assert hasattr(m["means2d"], "absgrad")
    assert (m["means2d"].absgrad + 1e-12 >= m["means2d"].grad.abs() * (1 - 1e-4)).all()


def test_golden_fixture(b2s):
    """CUDA path against the committed oracle outputs (tests/golden/scene_small.npz)."""
    from tests.golden.make_golden import SPEC
    gold = np.load(os.path.join(HERE, "golden", "scene_small.npz"))
    a = scene_args(SPEC["N"], SPEC["sh_degree"], SPEC["C"], SPEC["width"], SPEC["height"], seed=SPEC["seed"])
    g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}
    r, al, m = call(b2s.rasterization, g, SPEC["width"], SPEC["height"], sh_degree=SPEC["sh_degree"], packed=False,
                    rasterize_mode=SPEC["rasterize_mode"])
    for k in ("radii", "tiles_per_gauss", "isect_ids", "flatten_ids", "isect_offsets"):
        assert np.array_equal(m[k].cpu().numpy(), gold[k]), k
    fragile = torch.from_numpy(gold["fragile"])
    _compare_render(r.detach(), al.detach(), torch.from_numpy(gold["render"]), torch.from_numpy(gold["alpha"]), fragile)
    ((r * torch.from_numpy(gold["v_render"]).cuda()).sum() + (al * torch.from_numpy(gold["v_alpha"]).cuda()).sum()).backward()
    for k in LEAVES:
        rel, bad = grad_report(g[k].grad, torch.from_numpy(gold["grad_" + k]))
        assert rel < 1e-3 and bad < 5e-3, (k, rel, bad)


def test_edge_cases(b2s):
    # nothing visible: every Gaussian behind the camera
    W, H = 40, 24
    a = scene_args(64, 0, 1, W, H, seed=0)
    a["means"] = a["means"] * 0 + torch.tensor([0.0, 0.0, 50.0])
    g = {k: v.cuda() for k, v in a.items()}
    g["viewmats"] = torch.eye(4)[None].cuda()
    g["means"] = -g["means"]
    r, al, m = call(b2s.rasterization, g, W, H, sh_degree=0, packed=True)
    assert r.abs().max() == 0 and al.abs().max() == 0 and m["isect_ids"].numel() == 0 and m["means2d"].shape[0] == 0
    # a single Gaussian, image smaller than one tile
    one = {"means": torch.tensor([[0.0, 0.0, 3.0]]), "quats": torch.tensor([[1.0, 0, 0, 0]]), "scales": torch.full((1, 3), 0.3),
           "opacities": torch.tensor([0.7]), "colors": torch.tensor([[0.1, 0.6, 0.9]]), "viewmats": torch.eye(4)[None],
           "Ks": torch.tensor([[[20.0, 0, 5], [0, 20.0, 4], [0, 0, 1]]])}
    ref_r, ref_a, _ = call(O.rasterization, one, 10, 8, packed=False)
    got_r, got_a, _ = call(b2s.rasterization, {k: v.cuda() for k, v in one.items()}, 10, 8, packed=False)
    assert (got_r.cpu() - ref_r).abs().max() < 1e-5 and (got_a.cpu() - ref_a).abs().max() < 1e-5
    # unsupported options raise instead of falling back
    with pytest.raises(NotImplementedError):
        call(b2s.rasterization, g, W, H, camera_model="fisheye")
    with pytest.raises(NotImplementedError):
        call(b2s.rasterization, g, W, H, with_ut=True)
    with pytest.raises(Exception):
        call(b2s.rasterization, {k: v.cpu() for k, v in g.items()}, W, H)   # no CPU path


def test_full_size_coarse_bins_match_tile_lists(b2s, monkeypatch):
    """Config-3 size: two-level binning (lists expanded inside the compositing kernels) renders bit for bit what the
    reference-shaped pipeline (every tile intersection materialised and sorted) renders."""
    N, W, H = 5_800_000, 1297, 840
    a = scene_args(N, 3, 1, W, H, seed=0)
    g = {k: v.cuda() for k, v in a.items()}
    with torch.no_grad():
        monkeypatch.setenv("B2S_BINNING", "tiles")
        r, al, _ = call(b2s.rasterization, g, W, H, sh_degree=3, packed=False, rasterize_mode="antialiased")
        monkeypatch.setenv("B2S_BINNING", "coarse")
        rc, alc, _ = call(b2s.rasterization, g, W, H, sh_degree=3, packed=False, rasterize_mode="antialiased")
    assert torch.equal(r, rc) and torch.equal(al, alc)


def test_full_size_properties(b2s):
    """Config-3 size (5.8 M Gaussians, 1297x840, antialiased): properties that need no oracle."""
    N, W, H = 5_800_000, 1297, 840
    a = scene_args(N, 3, 1, W, H, seed=0)
    g = {k: v.cuda() for k, v in a.items()}
    with torch.no_grad():
        r, al, m = call(b2s.rasterization, g, W, H, sh_degree=3, packed=False, rasterize_mode="antialiased")
        ids, flat, off = m["isect_ids"], m["flatten_ids"], m["isect_offsets"].reshape(-1)
        assert int(m["tiles_per_gauss"].sum(dtype=torch.int64)) == ids.numel() == flat.numel()
        assert bool((ids[1:] >= ids[:-1]).all())                                  # sorted
        same = ids[1:] == ids[:-1]
        assert bool((flat[1:][same] > flat[:-1][same]).all())                     # stable
        assert bool((off[1:] >= off[:-1]).all()) and int(off[-1]) <= ids.numel()  # offsets monotone
        tile_of = (ids >> 32) & ((1 << 13) - 1)
        assert bool((tile_of[off[1:].long().clamp_max(ids.numel() - 1)] >= 0).all())
        # every key's depth bits belong to the Gaussian it names
        assert torch.equal((ids & 0xFFFFFFFF).int(), m["depths"].reshape(-1)[flat.long()].view(torch.int32))
        assert float(al.min()) >= 0 and float(al.max()) <= 1 and bool(torch.isfinite(r).all())
        rp, alp, mp = call(b2s.rasterization, g, W, H, sh_degree=3, packed=True, rasterize_mode="antialiased")
        assert torch.equal(r, rp) and torch.equal(al, alp)                        # packed == dense, bit for bit
        assert mp["means2d"].shape[0] == int((m["radii"] > 0).sum())


# ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,deg", [(1, 3), (2, 3), (1, 1), (1, None)])
def test_sparse_backward_equals_dense_backward(b2s, monkeypatch, C, deg):
    """A saturated scene: most visible rows come back with all-zero cotangents, so the backward of stage 1 switches (on
    the device) to the zero fill + list-driven kernels; the gradients must equal the dense kernels' (B2S_SPARSE_BWD=0)."""
    N, W, H = 60000, 96, 64
    a = scene_args(N, 3, C, W, H, seed=17)
    a["scales"] *= 3.0
    a["opacities"] = a["opacities"].clamp_min(0.6)
    gen = torch.Generator().manual_seed(2)
    v_r, v_a = torch.rand(C, H, W, 3, generator=gen).cuda(), torch.rand(C, H, W, 1, generator=gen).cuda()
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B2S_SPARSE_BWD", mode)
        g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}
        cols = g["colors"] if deg is not None else g["colors"][:, 0, :].detach().clone().requires_grad_()
        r, al, m = b2s.rasterization(g["means"], g["quats"], g["scales"], g["opacities"], cols, g["viewmats"], g["Ks"], W, H,
                                     sh_degree=deg, packed=False, rasterize_mode="antialiased")
        m["opacities"].retain_grad()
        torch.autograd.backward([r, al], [v_r, v_a])
        live = (m["opacities"].grad != 0).float().sum().item() / N
        assert live < 0.25, live                                  # the sparse mode really is the one taken
        out[mode] = [g[k].grad for k in ("means", "quats", "scales", "opacities")] + [cols.grad]
    for name, x, y in zip(("means", "quats", "scales", "opacities", "colors"), out["1"], out["0"]):
        assert x.shape == y.shape
        assert ((x == 0) == (y == 0)).all(), name                 # the same rows are touched
        assert grad_report(x, y)[0] < 1e-5, name


@pytest.mark.parametrize("N,C", [(70_001, 1), (5000, 3), (255, 1)])
def test_tma_coefficient_path_equals_ldg_path(b2s, monkeypatch, N, C):
    """Stage 1 with the degree-3 coefficient rows fetched by the tensor-memory accelerator (swizzled boxes, mbarrier) must
    produce bit for bit what the LDG/STS kernel produces (same arithmetic order), tail CTA included."""
    from pipeline_pointcloud_b200 import ops
    W, H = 160, 96
    a = scene_args(N, 3, C, W, H, seed=23, stress=True)
    g = {k: v.cuda() for k, v in a.items()}
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B2S_PROJECT_TMA", mode)
        o = ops._ProjectFused.apply(g["means"], g["quats"], g["scales"], g["opacities"], g["colors"], g["viewmats"], g["Ks"], None,
                                    W, H, 0.3, 0.01, 1e10, 0.0, True, False, 3, 16)
        torch.cuda.synchronize()
        outs[mode] = [t.clone() for t in o]
    assert (outs["1"][0] > 0).sum() > N // 4
    for x, y in zip(outs["1"], outs["0"]):
        assert torch.equal(x, y)
    # and against the fp64 oracle
    radii = outs["1"][0].cpu()
    campos = torch.linalg.inv(a["viewmats"])[:, :3, 3]
    dirs = a["means"][None] - campos[:, None]
    ref = torch.clamp_min(O.spherical_harmonics(3, dirs.double(), a["colors"][None].expand(C, N, 16, 3).double(), radii > 0) + 0.5, 0.0)
    ref = ref * (radii > 0)[..., None]
    assert (outs["1"][6].cpu().double() - ref).abs().max() < 2e-5


@pytest.mark.parametrize("mode,deg,C", [("antialiased", 3, 2), ("classic", 1, 1), ("classic", None, 3)])
def test_viewmat_gradients(b2s, mode, deg, C):
    """gsplat's `v_viewmats` (nerfstudio's camera optimizer reads it): gradients of the render w.r.t. the world->camera
    matrices -- through the means, the covariances and the SH view directions -- against the oracle's autograd."""
    N, W, H = 3000, 96, 64
    a = scene_args(N, deg if deg is not None else 0, C, W, H, seed=14)
    if deg is None:
        a["colors"] = torch.rand(N, 3, generator=torch.Generator().manual_seed(2))
    kw = dict(sh_degree=deg, packed=False, rasterize_mode=mode)
    ref_in = {k: v.clone() for k, v in a.items()}
    ref_in["viewmats"].requires_grad_()
    ref_in["means"].requires_grad_()
    ref_r, ref_a, ref_m = call(O.rasterization, ref_in, W, H, return_fragile=True, **kw)
    keep = (~ref_m["fragile"])[..., None].float()
    gen = torch.Generator().manual_seed(7)
    v_r, v_a = torch.rand(ref_r.shape, generator=gen) * keep, torch.rand(ref_a.shape, generator=gen) * keep
    ((ref_r * v_r).sum() + (ref_a * v_a).sum()).backward()
    g = {k: v.cuda() for k, v in a.items()}
    g["viewmats"].requires_grad_()
    g["means"].requires_grad_()
    r, al, m = call(b2s.rasterization, g, W, H, **kw)
    ((r * v_r.cuda()).sum() + (al * v_a.cuda()).sum()).backward()
    got, ref = g["viewmats"].grad.cpu().double(), ref_in["viewmats"].grad.double()
    assert got.shape == (C, 4, 4) and torch.all(got[:, 3] == 0)
    for c in range(C):
        rel = ((got[c, :3] - ref[c, :3]).norm() / ref[c, :3].norm()).item()
        assert rel < 1e-3, (c, rel, got[c], ref[c])
    assert grad_report(g["means"].grad, ref_in["means"].grad)[0] < 1e-3
    assert g["Ks"].grad is None


def test_packed_meta_read_does_not_keep_steps_alive(b2s):
    """The trainer of the reference reads ``meta["means2d"].grad`` of a ``packed=True`` call every iteration.  The packed rows of
    a step must die with the step: with the cyclic collector switched off, allocated device memory may not grow from step to step
    (a gradient hook that held the tensors strongly once kept every step's rows alive until a collection: 340 MB per step at
    config 3 and a cudaMalloc stall every few steps)."""
    import gc
    N, W, H = 40_000, 200, 120
    a = scene_args(N, 3, 1, W, H, seed=4)
    g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}

    def step():
        for k in LEAVES:
            g[k].grad = None
        r, al, m = call(b2s.rasterization, g, W, H, sh_degree=3, packed=True, absgrad=True)
        m["means2d"].retain_grad()
        (r.sum() + al.sum()).backward()
        assert m["means2d"].grad is not None and m["means2d"].absgrad is not None
        assert m["means2d"].grad.shape == m["means2d"].shape == (m["gaussian_ids"].shape[0], 2)

    gc.collect()
    gc.disable()
    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        for _ in range(8):
            step()
        torch.cuda.synchronize()
        grown = torch.cuda.memory_allocated() - base
    finally:
        gc.enable()
    assert grown < (1 << 20), f"allocated device memory grew by {grown} bytes over 8 steps"
