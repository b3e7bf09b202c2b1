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


def _worker(rank, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
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
        r, al, m = b2s.rasterization(shard["means"], shard["quats"], shard["scales"], shard["opacities"], shard["colors"],
                                     a["viewmats"][rank:rank + 1].to(dev), a["Ks"][rank:rank + 1].to(dev), W, H, sh_degree=DEG,
                                     packed=True, rasterize_mode="antialiased", distributed=True)
        m["means2d"].retain_grad()
        torch.autograd.backward([r, al], [v_r[rank:rank + 1].to(dev), v_a[rank:rank + 1].to(dev)])
        assert m["means2d"].grad is not None
        assert int(m["gaussian_ids"].max()) < N and int(m["camera_ids"].max()) == 0
        torch.save({"render": r.detach().cpu(), "alpha": al.detach().cpu(), "grads": {k: shard[k].grad.cpu() for k in LEAVES},
                    "n_rows": m["means2d"].shape[0]}, os.path.join(out, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_all_to_all_matches_single_gpu(tmp_path):
    import pipeline_pointcloud_b200 as b2s
    from tests.util import LEAVES, grad_report, scene_args
    mp.spawn(_worker, args=(29700, str(tmp_path)), nprocs=WORLD, join=True)
    a = scene_args(N, DEG, WORLD, W, H, seed=21)
    full = {k: (a[k].cuda().requires_grad_() if k in LEAVES else a[k].cuda()) for k in a}
    r, al, m = b2s.rasterization(full["means"], full["quats"], full["scales"], full["opacities"], full["colors"],
                                 full["viewmats"], full["Ks"], W, H, sh_degree=DEG, packed=True, rasterize_mode="antialiased")
    g = torch.Generator().manual_seed(5)
    v_r = torch.rand(WORLD, H, W, 3, generator=g)
    v_a = torch.rand(WORLD, H, W, 1, generator=g)
    torch.autograd.backward([r, al], [v_r.cuda(), v_a.cuda()])
    rows = 0
    for rank in range(WORLD):
        got = torch.load(os.path.join(str(tmp_path), f"rank{rank}.pt"))
        # same splats, same depth order: the composite differs only by the packed row numbering
        assert (got["render"][0] - r[rank].detach().cpu()).abs().max() < 2e-5
        assert (got["alpha"][0] - al[rank].detach().cpu()).abs().max() < 2e-5
        lo, hi = rank * N // WORLD, (rank + 1) * N // WORLD
        for k in LEAVES:      # each rank holds the complete gradient of its own shard (sum over both cameras)
            rel, bad = grad_report(got["grads"][k], full[k].grad[lo:hi])
            assert rel < 1e-4, (rank, k, rel)
        rows += got["n_rows"]
    assert rows == m["means2d"].shape[0]
