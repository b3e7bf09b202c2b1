"""GPU tests of the tile-row window / tile-stripe path (BASELINE config 5).

One GPU: binning and compositing restricted to a window of tile rows must give exactly the whole-image sorted lists
filtered by tile row (checked against the CPU oracle too) and exactly the window's pixel rows of the whole-image render.
Two GPUs (skipped with fewer): rasterization(distributed=True, tile_stripes=True) -- Gaussians sharded, every rank owns a
stripe of tile rows of the same camera -- against the single-GPU result."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import torch_oracle as O
from tests.util import LEAVES, grad_report, scene_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2s():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pipeline_pointcloud_b200 as pkg
    return pkg


WINDOWS = [(0, 4), (4, 10), (3, 3), (9, 10), (0, 10)]


@pytest.mark.parametrize("packed", [False, True])
def test_window_lists_are_the_filtered_whole_image_lists(b2s, packed):
    from pipeline_pointcloud_b200 import ops
    N, C, W, H = 20000, 2, 203, 150
    tw, th = (W + 15) // 16, (H + 15) // 16
    a = scene_args(N, 0, C, W, H, seed=31, stress=True)
    radii, means2d, depths, conics, _ = O.fully_fused_projection(a["means"], a["quats"], a["scales"], a["viewmats"], a["Ks"], W, H)
    kw = {}
    if packed:
        sel = (radii.reshape(-1) > 0).nonzero()[:, 0]
        radii, means2d, depths = radii.reshape(-1)[sel], means2d.reshape(-1, 2)[sel], depths.reshape(-1)[sel]
        kw = dict(packed=True, n_cameras=C, camera_ids=sel // N, gaussian_ids=sel % N)
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, tw, th, **kw)          # CPU oracle, whole image
    tile_bits = (tw * th).bit_length()
    trow = ((ids >> 32) & ((1 << tile_bits) - 1)) // tw
    kwg = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in kw.items()}
    m2, r, d = means2d.cuda(), radii.cuda(), depths.cuda()
    for lo, hi in WINDOWS:
        keep = (trow >= lo) & (trow < hi)
        for method in ("radix64", "depth_first"):
            wt, wids, wflat = b2s.isect_tiles(m2, r, d, 16, tw, th, method=method, tile_rows=(lo, hi), **kwg)
            assert torch.equal(wids.cpu(), ids[keep]), (lo, hi, method)
            assert torch.equal(wflat.cpu(), flat[keep]), (lo, hi, method)
            assert int(wt.sum()) == int(keep.sum())
        f2, off2, mk = ops.bin_depth_first(m2, r, d, wt, 16, tw, th, C, camera_ids=kwg.get("camera_ids"), tile_rows=(lo, hi))
        assert torch.equal(f2.cpu(), flat[keep]) and torch.equal(mk().cpu(), ids[keep])
        assert torch.equal(off2.cpu(), O.isect_offset_encode(ids[keep], C, tw, th))


@pytest.mark.parametrize("use_bins", [True, False])
def test_window_render_is_the_rows_of_the_whole_render(b2s, use_bins):
    from pipeline_pointcloud_b200 import ops
    N, C, W, H = 12000, 2, 203, 150
    tw, th = (W + 15) // 16, (H + 15) // 16
    a = scene_args(N, 0, C, W, H, seed=32)
    g = {k: v.cuda() for k, v in a.items()}
    radii, means2d, depths, conics, _ = b2s.fully_fused_projection(g["means"], None, g["quats"], g["scales"], g["viewmats"],
                                                                   g["Ks"], W, H)
    opac = g["opacities"][None].expand(C, N).contiguous()
    cols = torch.rand(C, N, 3, device="cuda")
    bg = torch.rand(C, 3, device="cuda")
    gen = torch.Generator().manual_seed(3)
    v_r, v_a = torch.rand(C, H, W, 3, generator=gen).cuda(), torch.rand(C, H, W, 1, generator=gen).cuda()

    def run(rows):
        leaves = [t.clone().requires_grad_() for t in (means2d, conics, cols, opac)]
        if use_bins:
            bins = ops.bin_coarse(means2d, radii, depths, 16, tw, th, C, tile_rows=rows)
            out = ops.rasterize_to_pixels(leaves[0], leaves[1], leaves[2], leaves[3], W, H, 16, None, None, backgrounds=bg,
                                          bins=bins, n_cameras=C, tile_rows=rows)
        else:
            tpg = b2s.isect_tiles(means2d, radii, depths, 16, tw, th, tile_rows=rows, count_only=True)
            flat, off, _ = ops.bin_depth_first(means2d, radii, depths, tpg, 16, tw, th, C, tile_rows=rows)
            out = ops.rasterize_to_pixels(leaves[0], leaves[1], leaves[2], leaves[3], W, H, 16, off, flat, backgrounds=bg,
                                          tile_rows=rows)
        y0, y1 = (0, H) if rows is None else (rows[0] * 16, min(H, rows[1] * 16))
        if y1 > y0:
            torch.autograd.backward(list(out), [v_r[:, y0:y1].contiguous(), v_a[:, y0:y1].contiguous()])
        return out, [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves], (y0, y1)

    (full_r, full_a), full_g, _ = run(None)
    acc = [torch.zeros_like(t) for t in full_g]
    for rows in [(0, 4), (4, 4), (4, 9), (9, 10)]:          # a partition of the 10 tile rows (one stripe empty)
        (r, al), gr, (y0, y1) = run(rows)
        assert r.shape == (C, y1 - y0, W, 3) and al.shape == (C, y1 - y0, W, 1)
        assert torch.equal(r, full_r[:, y0:y1]) and torch.equal(al, full_a[:, y0:y1])      # bit for bit
        acc = [x + y for x, y in zip(acc, gr)]
    for name, x, y in zip(("means2d", "conics", "colors", "opacities"), acc, full_g):
        assert grad_report(x, y)[0] < 1e-5, name


# ----------------------------------------------------------------------------------------------------------
# 2 GPUs: the whole stripe path
# ----------------------------------------------------------------------------------------------------------
N2, DEG, W2, H2, C2 = 8000, 3, 176, 150, 2
TH2 = (H2 + 15) // 16          # 10 tile rows
# (world, stripe bounds): even-ish, uneven, and with EMPTY stripes (a rank that owns no tile row still projects its shard,
# takes part in the exchange and gets the gradients of its Gaussians back)
STRIPE_CASES = {
    "w2": (2, [0, 4, TH2]),
    "w2_empty_first": (2, [0, 0, TH2]),
    "w4_uneven_one_empty": (4, [0, 3, 3, 7, TH2]),
    "w8_uneven_one_empty": (8, [0, 1, 3, 3, 4, 6, 7, 9, TH2]),
}


def _worker(rank, world, bounds, port, out, route):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["B2S_ROUTE"] = route
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import pipeline_pointcloud_b200 as b2s
        from pipeline_pointcloud_b200.distributed import gather_stripes
        a = scene_args(N2, DEG, C2, W2, H2, seed=41)
        a["scales"][:200] *= 5.0                      # splats that straddle the stripe boundaries
        dev = torch.device("cuda", rank)
        lo, hi = rank * N2 // world, (rank + 1) * N2 // world
        shard = {k: a[k][lo:hi].to(dev).requires_grad_() for k in LEAVES}
        g = torch.Generator().manual_seed(5)
        v_r = torch.rand(C2, H2, W2, 3, generator=g).to(dev)
        v_a = torch.rand(C2, H2, W2, 1, generator=g).to(dev)
        for it in range(2):                           # twice: the second pass reuses the exchange buffers and the table
            for t in shard.values():
                t.grad = None
            r, al, m = b2s.rasterization(shard["means"], shard["quats"], shard["scales"], shard["opacities"], shard["colors"],
                                         a["viewmats"].to(dev), a["Ks"].to(dev), W2, H2, sh_degree=DEG, packed=True,
                                         rasterize_mode="antialiased", distributed=True, tile_stripes=True, stripe_rows=bounds)
            y0, y1 = m["stripe_pixel_rows"]
            assert r.shape == (C2, y1 - y0, W2, 3)
            full_r, full_a = gather_stripes(r, m), gather_stripes(al, m)
            assert full_r.shape == (C2, H2, W2, 3)
            torch.autograd.backward([full_r, full_a], [v_r, v_a])      # every rank evaluates the same "loss" on the full image
        torch.save({"render": full_r.detach().cpu(), "alpha": full_a.detach().cpu(),
                    "grads": {k: shard[k].grad.cpu() for k in LEAVES}, "isect_ids": m["isect_ids"].cpu(),
                    "gids": m["gaussian_ids"][m["flatten_ids"].long()].cpu(), "cams": m["camera_ids"][m["flatten_ids"].long()].cpu()},
                   os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case,route", [("w2", "direct"), ("w2", "p2p"), ("w2", "nccl"), ("w2_empty_first", "direct"),
                                        ("w2_empty_first", "p2p"), ("w2_empty_first", "nccl"), ("w4_uneven_one_empty", "direct"),
                                        ("w4_uneven_one_empty", "p2p"), ("w4_uneven_one_empty", "nccl"),
                                        ("w8_uneven_one_empty", "direct")])
def test_tile_stripes_match_single_gpu(b2s, tmp_path, case, route):
    world, bounds = STRIPE_CASES[case]
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29720 + 7 * list(STRIPE_CASES).index(case) + ("direct", "p2p", "nccl").index(route)
    mp.spawn(_worker, args=(world, bounds, port, str(tmp_path), route), nprocs=world, join=True)
    a = scene_args(N2, DEG, C2, W2, H2, seed=41)
    a["scales"][:200] *= 5.0
    full = {k: (a[k].cuda().requires_grad_() if k in LEAVES else a[k].cuda()) for k in a}
    r, al, m = b2s.rasterization(full["means"], full["quats"], full["scales"], full["opacities"], full["colors"],
                                 full["viewmats"], full["Ks"], W2, H2, sh_degree=DEG, packed=True, rasterize_mode="antialiased")
    g = torch.Generator().manual_seed(5)
    v_r = torch.rand(C2, H2, W2, 3, generator=g)
    v_a = torch.rand(C2, H2, W2, 1, generator=g)
    torch.autograd.backward([r, al], [v_r.cuda(), v_a.cuda()])
    tw, th = (W2 + 15) // 16, TH2
    ids = m["isect_ids"].cpu()
    gids = m["gaussian_ids"][m["flatten_ids"].long()].cpu()
    trow = ((ids >> 32) & ((1 << (tw * th).bit_length()) - 1)) // tw
    total = 0
    for rank in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        # same per-tile lists in the same order => the very same arithmetic per pixel
        assert torch.equal(got["render"], r.detach().cpu()) and torch.equal(got["alpha"], al.detach().cpu())
        # the stripe's sorted key list = the single-GPU list filtered by tile row, same Gaussians in the same order
        keep = (trow >= bounds[rank]) & (trow < bounds[rank + 1])
        assert torch.equal(got["isect_ids"], ids[keep])
        assert torch.equal(got["gids"], gids[keep])
        total += int(keep.sum())
        lo, hi = rank * N2 // world, (rank + 1) * N2 // world
        for k in LEAVES:      # each rank holds the complete gradient of its own shard (summed over stripes and cameras)
            rel, bad = grad_report(got["grads"][k], full[k].grad[lo:hi])
            assert rel < 1e-4, (rank, k, rel)
    assert total == ids.numel()                     # the stripes partition the single-GPU list
