"""CPU (gloo, world_size 2) tests of the multi-GPU host logic: the all-to-all routing of projected splats to the
camera owners, its reverse in the backward, and the data-parallel gradient all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pipeline_pointcloud_b200 import distributed as D

WORLD = 2


def _init(rank, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)


def _fake_rows(rank, cams_per_rank, n_local):
    """Deterministic packed rows of rank `rank`: visible (camera, gaussian) pairs, camera-major."""
    g = torch.Generator().manual_seed(100 + rank)
    C_all = cams_per_rank * WORLD
    vis = torch.rand(C_all, n_local, generator=g) > 0.4
    sel = vis.reshape(-1).nonzero()[:, 0]
    cam, gid = sel // n_local, sel % n_local
    # payload encodes (source rank, camera, gaussian) so the receiver can check what it got
    payload = torch.stack([torch.full_like(cam, rank), cam, gid], 1).double() + 0.25
    ids = torch.stack([cam * 1000 + gid, gid * 7 + rank], 1)
    return payload, ids, cam, gid


def _route_worker(rank, port, cams_per_rank, n_locals):
    _init(rank, port)
    try:
        payload, ids, cam, gid = _fake_rows(rank, cams_per_rank, n_locals[rank])
        payload.requires_grad_()
        out, oids, cam_local, g_global, sc, rc = D.route_to_camera_owners(payload, ids, cam, gid, cams_per_rank, n_locals)
        # expected: rows of every source rank whose camera belongs to me, in source-rank order
        exp_rows, exp_ids, exp_cam, exp_g = [], [], [], []
        for src in range(WORLD):
            p, i, c, g = _fake_rows(src, cams_per_rank, n_locals[src])
            m = (c // cams_per_rank) == rank
            exp_rows.append(p[m]); exp_ids.append(i[m]); exp_cam.append(c[m] - rank * cams_per_rank)
            exp_g.append(g[m] + sum(n_locals[:src]))
        assert torch.equal(out.detach(), torch.cat(exp_rows))
        assert torch.equal(oids, torch.cat(exp_ids))
        assert torch.equal(cam_local, torch.cat(exp_cam)) and torch.equal(g_global, torch.cat(exp_g))
        assert (cam_local >= 0).all() and (cam_local < cams_per_rank).all()
        assert sum(rc) == out.shape[0] and sum(sc) == payload.shape[0]
        # backward: every source row gets back exactly the cotangent its owner attached to it
        w = out.detach()[:, 1] * 10 + out.detach()[:, 2] + 1000 * rank       # depends on (camera, gaussian, owner)
        (out[:, 0] * w).sum().backward()
        owner = (cam // cams_per_rank).double()
        assert torch.allclose(payload.grad[:, 0], (cam.double() + 0.25) * 10 + (gid.double() + 0.25) + 1000 * owner)
        assert payload.grad[:, 1:].abs().sum() == 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cams_per_rank,n_locals", [(1, [50, 37]), (3, [20, 64])])
def test_route_to_camera_owners(cams_per_rank, n_locals):
    port = 29600 + cams_per_rank
    mp.spawn(_route_worker, args=(port, cams_per_rank, n_locals), nprocs=WORLD, join=True)


def _allreduce_worker(rank, port):
    _init(rank, port)
    try:
        ps = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(40))]
        for i, p in enumerate(ps):
            p.grad = torch.full_like(p, float(rank + 1 + i))
        D.allreduce_gradients(ps)
        assert torch.equal(ps[0].grad, torch.full((5, 3), 3.0)) and torch.equal(ps[1].grad, torch.full((40,), 5.0))
        vm, ks, ns = D.all_gather_cameras(torch.full((2, 4, 4), float(rank)), torch.full((2, 3, 3), float(rank)), 10 + rank)
        assert vm.shape == (4, 4, 4) and ks.shape == (4, 3, 3) and ns == [10, 11]
        assert vm[:2].eq(0).all() and vm[2:].eq(1).all()
        assert D.cameras_of_rank(8, rank, WORLD) == list(range(rank, 8, WORLD))
    finally:
        dist.destroy_process_group()


def test_allreduce_and_camera_gather():
    mp.spawn(_allreduce_worker, args=(29650,), nprocs=WORLD, join=True)


# ----------------------------------------------------------------------------------------------------------
# tile stripes: Gaussian-sharded projection, rows routed to the owners of tile-row stripes
# ----------------------------------------------------------------------------------------------------------
S_N, S_C, S_W, S_H = 900, 2, 200, 150          # th = 10 tile rows


def _stripe_scene():
    from oracle import torch_oracle as O           # the checker of the routing decision
    from pipeline_pointcloud_b200.scene import make_scene
    s = make_scene(S_N, 0, S_C, S_W, S_H, seed=11)
    s["scales"][:60] *= 6.0                        # some splats that straddle stripe boundaries
    radii, means2d, depths, conics, _ = O.fully_fused_projection(s["means"], s["quats"], s["scales"], s["viewmats"], s["Ks"],
                                                                 S_W, S_H)
    return O, radii, means2d, depths


def test_stripe_selection_matches_oracle_tile_rects():
    O, radii, means2d, depths = _stripe_scene()
    th = (S_H + 15) // 16
    tmin, tmax = O.tile_rects(means2d.reshape(-1, 2), radii.reshape(-1), 16, (S_W + 15) // 16, th)
    for bounds in ([0, 5, 10], [0, 3, 3, 7, 10], [0, 10]):
        sel, counts = D.stripe_selection(means2d, radii, 16, th, bounds)
        start = 0
        for d in range(len(bounds) - 1):
            lo, hi = bounds[d], bounds[d + 1]
            y0, y1 = tmin[:, 1].clamp(lo, hi), tmax[:, 1].clamp(lo, hi)
            exp = ((radii.reshape(-1) > 0) & (y1 > y0)).nonzero()[:, 0]
            got = sel[start:start + int(counts[d])]
            start += int(counts[d])
            assert torch.equal(got, exp), (bounds, d)
        assert start == sel.numel()
    assert D.stripe_bounds(135, 8) == [0, 16, 33, 50, 67, 84, 101, 118, 135]
    assert D.balanced_stripe_bounds(torch.tensor([1, 1, 1, 1, 10, 10, 1, 1]), 2) == [0, 5, 8]
    b = D.balanced_stripe_bounds(torch.zeros(9), 3)
    assert b == D.stripe_bounds(9, 3)
    # feedback balancing: cost concentrated in the middle stripe -> the middle stripe shrinks, the cut points stay ordered
    nb = D.rebalance_stripe_bounds([0, 40, 80, 120], [1.0, 4.0, 1.0])
    assert nb[0] == 0 and nb[-1] == 120 and nb == sorted(nb) and nb[2] - nb[1] < 40 and nb[1] > 40 and nb[2] < 80
    assert D.rebalance_stripe_bounds([0, 40, 80, 120], [1.0, 1.0, 1.0]) == [0, 40, 80, 120]
    assert D.rebalance_stripe_bounds([0, 1, 2, 3], [5.0, 0.0, 0.0]) == [0, 1, 2, 3]          # nothing left to move
    assert D.rebalance_stripe_bounds([0, 10], [3.0]) == [0, 10]


def _stripe_worker(rank, port):
    _init(rank, port)
    try:
        O, radii, means2d, depths = _stripe_scene()
        th = (S_H + 15) // 16
        bounds = [0, 4, th]
        n_locals = [S_N // 3, S_N - S_N // 3]
        lo_n = sum(n_locals[:rank])
        sl = slice(lo_n, lo_n + n_locals[rank])
        r_l, m_l = radii[:, sl].contiguous(), means2d[:, sl].contiguous()
        Nl = n_locals[rank]
        cam = torch.arange(S_C)[:, None].expand(S_C, Nl)
        gid = torch.arange(Nl)[None, :].expand(S_C, Nl) + lo_n
        payload = torch.stack([cam, gid], -1).reshape(-1, 2).double().requires_grad_()
        ids = (cam * 10000 + gid).reshape(-1, 1)
        out, oids, cams, gids, sc, rc = D.route_to_stripe_owners(payload, ids, m_l, r_l, 16, th, bounds, n_locals)
        # expected: every (camera, Gaussian) of the whole scene whose tile rows touch my stripe, source-rank-major,
        # camera-major inside a source, Gaussian-ascending inside a camera
        tmin, tmax = O.tile_rects(means2d.reshape(-1, 2), radii.reshape(-1), 16, (S_W + 15) // 16, th)
        lo, hi = bounds[rank], bounds[rank + 1]
        hit = ((radii.reshape(-1) > 0) & (tmax[:, 1].clamp(lo, hi) > tmin[:, 1].clamp(lo, hi))).reshape(S_C, S_N)
        exp_c, exp_g = [], []
        for src in range(WORLD):
            s0 = sum(n_locals[:src])
            h = hit[:, s0:s0 + n_locals[src]]
            c, g = h.nonzero(as_tuple=True)
            exp_c.append(c); exp_g.append(g + s0)
        exp_c, exp_g = torch.cat(exp_c), torch.cat(exp_g)
        assert torch.equal(cams, exp_c) and torch.equal(gids, exp_g)
        assert torch.equal(out.detach()[:, 0].long(), exp_c) and torch.equal(out.detach()[:, 1].long(), exp_g)
        assert torch.equal(oids[:, 0], exp_c * 10000 + exp_g)
        # per camera, the received rows are in ascending global Gaussian id (the tie-break order of the stable sort)
        for c in range(S_C):
            g = gids[cams == c]
            assert (g[1:] > g[:-1]).all()
        # backward: a row sent to both stripes receives the sum of both owners' cotangents
        (out[:, 0] * (rank + 1)).sum().backward()
        tmin_l, tmax_l = tmin[:, 1].reshape(S_C, S_N)[:, sl].reshape(-1), tmax[:, 1].reshape(S_C, S_N)[:, sl].reshape(-1)
        vis = r_l.reshape(-1) > 0
        exp_grad = torch.zeros(S_C * Nl, dtype=torch.float64)
        for d in range(WORLD):
            dl, dh = bounds[d], bounds[d + 1]
            exp_grad += (vis & (tmax_l.clamp(dl, dh) > tmin_l.clamp(dl, dh))).double() * (d + 1)
        assert torch.equal(payload.grad[:, 0], exp_grad)
        assert (exp_grad == 3).sum() > 0, "the scene must contain splats that straddle the stripe boundary"
    finally:
        dist.destroy_process_group()


def test_route_to_stripe_owners():
    mp.spawn(_stripe_worker, args=(29670,), nprocs=WORLD, join=True)
