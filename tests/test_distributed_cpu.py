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
