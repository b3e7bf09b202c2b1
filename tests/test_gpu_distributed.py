"""Multi-GPU test (needs >= 2 GPUs; skipped otherwise): rasterization(distributed=True) on 2 ranks -- Gaussians
sharded, one camera per rank, NCCL all-to-all -- against the single-GPU render of the whole scene."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N, DEG, W, H = 6000, 3, 160, 96
WORLD = 2


def _worker(rank, port, out, route, render_mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["B2S_ROUTE"] = route
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        import pipeline_pointcloud_b200 as b2s
        from tests.util import LEAVES, scene_args
        a = scene_args(N, DEG, WORLD, W, H, seed=21)
        dev = torch.device("cuda", rank)
        lo, hi = rank * N // WORLD, (rank + 1) * N // WORLD
        shard = {k: a[k][lo:hi].to(dev).requires_grad_() for k in LEAVES}
        g = torch.Generator().manual_seed(5)
        v_r = torch.rand(WORLD, H, W, 3, generator=g)
        v_a = torch.rand(WORLD, H, W, 1, generator=g)
        nch = 4 if render_mode == "RGB+D" else 3
        v_r = torch.rand(WORLD, H, W, nch, generator=g)
        for it in range(2):                   # twice: the second pass alternates the exchange buffers
            for t in shard.values():
                t.grad = None
            r, al, m = b2s.rasterization(shard["means"], shard["quats"], shard["scales"], shard["opacities"], shard["colors"],
                                         a["viewmats"][rank:rank + 1].to(dev), a["Ks"][rank:rank + 1].to(dev), W, H, sh_degree=DEG,
                                         packed=True, rasterize_mode="antialiased", distributed=True, render_mode=render_mode)
            m["means2d"].retain_grad()
            torch.autograd.backward([r, al], [v_r[rank:rank + 1].to(dev), v_a[rank:rank + 1].to(dev)])
        assert m["means2d"].grad is not None
        assert int(m["gaussian_ids"].max()) < N and int(m["camera_ids"].max()) == 0
        torch.save({"render": r.detach().cpu(), "alpha": al.detach().cpu(), "grads": {k: shard[k].grad.cpu() for k in LEAVES},
                    "n_rows": m["means2d"].shape[0]}, os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("route,render_mode", [("direct", "RGB"), ("direct", "RGB+D"), ("p2p", "RGB"), ("nccl", "RGB")])
def test_sharded_all_to_all_matches_single_gpu(tmp_path, route, render_mode):
    import pipeline_pointcloud_b200 as b2s
    from tests.util import LEAVES, grad_report, scene_args
    mp.spawn(_worker, args=(29700 + ("direct", "p2p", "nccl").index(route) + 3 * (render_mode != "RGB"), str(tmp_path), route,
                            render_mode), nprocs=WORLD, join=True)
    a = scene_args(N, DEG, WORLD, W, H, seed=21)
    full = {k: (a[k].cuda().requires_grad_() if k in LEAVES else a[k].cuda()) for k in a}
    r, al, m = b2s.rasterization(full["means"], full["quats"], full["scales"], full["opacities"], full["colors"],
                                 full["viewmats"], full["Ks"], W, H, sh_degree=DEG, packed=True, rasterize_mode="antialiased",
                                 render_mode=render_mode)
    g = torch.Generator().manual_seed(5)
    v_r = torch.rand(WORLD, H, W, 3, generator=g)
    v_a = torch.rand(WORLD, H, W, 1, generator=g)
    v_r = torch.rand(WORLD, H, W, 4 if render_mode == "RGB+D" else 3, generator=g)
    torch.autograd.backward([r, al], [v_r.cuda(), v_a.cuda()])
    rows = 0
    for rank in range(WORLD):
        got = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        # same splats, same depth order: the composite differs only by the packed row numbering
        assert (got["render"][0][..., :3] - r[rank].detach().cpu()[..., :3]).abs().max() < 2e-5
        if render_mode == "RGB+D":
            assert (got["render"][0][..., 3] - r[rank].detach().cpu()[..., 3]).abs().max() < 2e-4
        assert (got["alpha"][0] - al[rank].detach().cpu()).abs().max() < 2e-5
        lo, hi = rank * N // WORLD, (rank + 1) * N // WORLD
        for k in LEAVES:      # each rank holds the complete gradient of its own shard (sum over both cameras)
            rel, bad = grad_report(got["grads"][k], full[k].grad[lo:hi])
            assert rel < 1e-4, (rank, k, rel)
        rows += got["n_rows"]
    assert rows == m["means2d"].shape[0]


def test_grad_compact_and_scatter_add_single_gpu():
    """The two kernels of the sparse gradient synchronisation (csrc/gradsync.cu) without a process group: the records of
    rank A's live rows, added into rank B's dense gradients, give A + B."""
    from pipeline_pointcloud_b200 import _lib as L
    torch.manual_seed(0)
    N, K = 20_000, 16
    dev = torch.device("cuda", 0)
    shapes = [(N, 3), (N, 4), (N, 3), (N,), (N, K, 3)]

    def sparse_grads(seed, frac):
        g = torch.Generator(device="cpu").manual_seed(seed)
        live = torch.rand(N, generator=g) < frac
        out = []
        for s in shapes:
            t = torch.randn(s, generator=g)
            t[~live] = 0
            out.append(t.to(dev))
        # a row whose only gradient is its colour: still live (first SH row non-zero)
        out[0][live.nonzero()[0, 0]] = 0; out[1][live.nonzero()[0, 0]] = 0; out[2][live.nonzero()[0, 0]] = 0; out[3][live.nonzero()[0, 0]] = 0
        return out, int(live.sum())

    a, na = sparse_grads(1, 0.03)
    b, _ = sparse_grads(2, 0.05)
    lib, st = L.lib(), L.stream_ptr(dev)
    R = 12 + 3 * K
    total = torch.empty(1, dtype=torch.int64, device=dev)
    rec = torch.empty(N // 8, R, device=dev)
    L.check(lib.b2s_grad_compact(N, 3 * K, *[L.ptr(t) for t in a], L.ptr(total), L.ptr(rec), N // 8, st), "compact")
    assert int(total) == na
    rows = rec[:na, 0].contiguous().view(torch.int32)
    assert rows.unique().numel() == na
    want = [x + y for x, y in zip(a, b)]
    L.check(lib.b2s_grad_scatter_add(na, 3 * K, L.ptr(rec), *[L.ptr(t) for t in b], st), "scatter")
    for x, y in zip(b, want):
        assert torch.equal(x, y)                       # one add per element: exact
    # capacity overflow is reported, not written past
    L.check(lib.b2s_grad_compact(N, 3 * K, *[L.ptr(t) for t in a], L.ptr(total), L.ptr(rec), 10, st), "compact")
    assert int(total) == na


def _allreduce_worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        import pipeline_pointcloud_b200 as b2s
        from pipeline_pointcloud_b200.distributed import allreduce_gradients, allreduce_splat_gradients
        from tests.util import LEAVES, scene_args
        Nn = 40_000
        a = scene_args(Nn, DEG, WORLD, W, H, seed=33)
        a["scales"] *= 3.0                       # saturating scene: few rows carry a gradient
        a["opacities"] = a["opacities"].clamp_min(0.6)
        dev = torch.device("cuda", rank)
        p = {k: a[k].to(dev).requires_grad_() for k in LEAVES}
        g = torch.Generator().manual_seed(5)
        v_r, v_a = torch.rand(WORLD, H, W, 3, generator=g), torch.rand(WORLD, H, W, 1, generator=g)
        r, al, m = b2s.rasterization(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"],
                                     a["viewmats"][rank:rank + 1].to(dev), a["Ks"][rank:rank + 1].to(dev), W, H, sh_degree=DEG,
                                     packed=False, rasterize_mode="antialiased")
        torch.autograd.backward([r, al], [v_r[rank:rank + 1].to(dev), v_a[rank:rank + 1].to(dev)])
        dense = {k: p[k].grad.clone() for k in LEAVES}
        for t in dense.values():
            dist.all_reduce(t)
        stats = {}
        mode = allreduce_splat_gradients(p["means"], p["quats"], p["scales"], p["opacities"], p["colors"], stats=stats)
        torch.save({"mode": mode, "live": stats.get("live_rows"), "sparse": {k: p[k].grad.cpu() for k in LEAVES},
                    "dense": {k: v.cpu() for k, v in dense.items()}}, os.path.join(out, f"ar{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sparse_gradient_allreduce_matches_dense(tmp_path):
    from tests.util import LEAVES, grad_report
    mp.spawn(_allreduce_worker, args=(29790, str(tmp_path)), nprocs=WORLD, join=True)
    for rank in range(WORLD):
        got = torch.load(os.path.join(str(tmp_path), f"ar{rank}.pt"))
        assert got["mode"] == "sparse" and max(got["live"]) < 40_000 // 8, got["live"]
        for k in LEAVES:
            assert grad_report(got["sparse"][k], got["dense"][k])[0] < 1e-6, k
            assert ((got["sparse"][k] == 0) == (got["dense"][k] == 0)).all(), k
