"""CPU tests that pin the oracle itself (the reference holds no golden vector for this path, so
the oracle is checked against hand-computable micro-scenes, self-consistency properties, fp64 finite
differences, and the committed golden file that it generated)."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from pipeline_pointcloud_b200.scene import make_scene

HERE = os.path.dirname(os.path.abspath(__file__))


def one_gaussian_scene(scale=0.2, z=4.0, W=64, H=48, f=50.0, opacity=0.8, color=(0.2, 0.5, 0.9), xy=(0.0, 0.0)):
    means = torch.tensor([[xy[0], xy[1], z]])
    quats = torch.tensor([[1.0, 0.0, 0.0, 0.0]])
    scales = torch.full((1, 3), scale)
    opac = torch.tensor([opacity])
    cols = torch.tensor([color])
    viewmats = torch.eye(4)[None]
    Ks = torch.tensor([[[f, 0.0, W / 2], [0.0, f, H / 2], [0.0, 0.0, 1.0]]])
    return means, quats, scales, opac, cols, viewmats, Ks, W, H


def test_single_isotropic_gaussian_closed_form():
    means, quats, scales, opac, cols, V, K, W, H = one_gaussian_scene()
    render, alpha, meta = O.rasterization(means, quats, scales, opac, cols, V, K, W, H, packed=False)
    f, z, s = 50.0, 4.0, 0.2
    var = (s * f / z) ** 2 + 0.3
    assert torch.allclose(meta["means2d"][0, 0], torch.tensor([W / 2, H / 2]))
    assert torch.allclose(meta["conics"][0, 0], torch.tensor([1 / var, 0.0, 1 / var]), atol=1e-6)
    assert meta["radii"][0, 0].item() == math.ceil(3 * math.sqrt(var))
    ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
    r2 = (xs - W / 2) ** 2 + (ys - H / 2) ** 2
    a = 0.8 * torch.exp(-0.5 * r2 / var)
    a = torch.where(a >= 1 / 255, a, torch.zeros_like(a))
    # only tiles inside the radius-bounding box are rasterised
    rad = meta["radii"][0, 0].item()
    tmin, tmax = O.tile_rects(meta["means2d"][0], meta["radii"][0], 16, 4, 3)
    mask = torch.zeros(H, W, dtype=torch.bool)
    mask[tmin[0, 1] * 16:tmax[0, 1] * 16, tmin[0, 0] * 16:tmax[0, 0] * 16] = True
    a = a * mask
    assert torch.allclose(alpha[0, ..., 0], a, atol=1e-6)
    assert torch.allclose(render[0], a[..., None] * cols[0], atol=1e-6)
    assert rad > 0


def test_front_to_back_order_and_early_stop():
    # two opaque-ish Gaussians on the same ray: the nearer one dominates; stacking many saturates T
    means = torch.tensor([[0.0, 0.0, 6.0], [0.0, 0.0, 3.0]])
    quats = torch.tensor([[1.0, 0, 0, 0]]).repeat(2, 1)
    scales = torch.full((2, 3), 0.5)
    opac = torch.tensor([0.9, 0.9])
    cols = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    V = torch.eye(4)[None]
    K = torch.tensor([[[40.0, 0, 16], [0, 40.0, 16], [0, 0, 1]]])
    render, alpha, meta = O.rasterization(means, quats, scales, opac, cols, V, K, 32, 32, packed=False)
    c = render[0, 16, 16]
    assert c[1] > 0.85 and c[0] < 0.1          # green (near) in front of red (far)
    assert meta["flatten_ids"][0].item() == 1    # nearer Gaussian sorts first in its tile
    n = 40
    means = torch.zeros(n, 3); means[:, 2] = torch.linspace(3, 8, n)
    render, alpha, meta = O.rasterization(means, quats[:1].repeat(n, 1), torch.full((n, 3), 0.5), torch.full((n,), 0.9),
                                          torch.rand(n, 3), V, K, 32, 32, packed=False)
    # T after k Gaussians = 0.1^k ; the one that would push T <= 1e-4 (k = 4) is not composited
    assert abs(alpha[0, 16, 16, 0].item() - (1 - 0.1 ** 3)) < 2e-3


def test_tile_corner_gives_four_keys():
    means, quats, scales, opac, cols, V, K, W, H = one_gaussian_scene(scale=0.02, xy=(0.0, 0.0), W=64, H=64, f=50.0)
    # centre (32,32) is the corner shared by tiles (1,1),(2,1),(1,2),(2,2)
    radii, means2d, depths, conics, _ = O.fully_fused_projection(means, quats, scales, V, K, W, H)
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, 4, 4)
    assert tpg.item() == 4
    tiles = ((ids >> 32) & 0x1F).tolist()
    assert tiles == [5, 6, 9, 10]
    assert ((ids & 0xFFFFFFFF) == depths.view(torch.int32).item()).all()
    off = O.isect_offset_encode(ids, 1, 4, 4).reshape(-1).tolist()
    assert off == [0, 0, 0, 0, 0, 0, 1, 2, 2, 2, 3, 4, 4, 4, 4, 4]


def test_conic_is_inverse_of_blurred_cov_and_counts_add_up():
    s = make_scene(4000, 1, 2, 160, 96, seed=3)
    radii, means2d, depths, conics, comps = O.fully_fused_projection(
        s["means"], s["quats"], s["scales"], s["viewmats"], s["Ks"], 160, 96, calc_compensations=True)
    vis = radii > 0
    a, b, c = conics[vis].unbind(-1)
    det = a * c - b * b
    cov = torch.stack([c / det, -b / det, a / det], -1)     # inverse of the conic = blurred covariance
    assert (cov[:, 0] >= 0.3 - 1e-4).all() and (cov[:, 2] >= 0.3 - 1e-4).all()
    assert ((comps[vis] > 0) & (comps[vis] <= 1 + 1e-6)).all()
    tpg, ids, flat = O.isect_tiles(means2d, radii, depths, 16, 10, 6)
    assert tpg.sum().item() == ids.numel() == flat.numel()
    assert (ids[1:] >= ids[:-1]).all()
    off = O.isect_offset_encode(ids, 2, 10, 6).reshape(-1)
    assert (off[1:] >= off[:-1]).all() and off[-1] <= ids.numel()
    # stable: equal keys keep ascending flat index
    same = ids[1:] == ids[:-1]
    assert (flat[1:][same] > flat[:-1][same]).all()


def test_sh_basis_is_orthonormal():
    g = torch.Generator().manual_seed(0)
    d = torch.randn(400000, 3, generator=g, dtype=torch.float64)
    d = d / d.norm(dim=-1, keepdim=True)
    Y = O.eval_sh_bases(3, d)
    gram = (Y.T @ Y) / d.shape[0] * 4 * math.pi
    assert (gram - torch.eye(16, dtype=torch.float64)).abs().max() < 0.02


def test_weights_sum_to_alpha_and_packed_equals_dense():
    s = make_scene(1500, 2, 2, 80, 64, seed=5)
    args = (s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmats"], s["Ks"], 80, 64)
    ones = torch.ones(1500, 3)
    r1, a1, _ = O.rasterization(s["means"], s["quats"], s["scales"], s["opacities"], ones, s["viewmats"], s["Ks"], 80, 64,
                                packed=False)
    assert torch.allclose(r1[..., 0:1], a1, atol=2e-6)    # colour == 1 everywhere => render == sum of weights == alpha
    rd, ad, md = O.rasterization(*args, sh_degree=2, packed=False, rasterize_mode="antialiased")
    rp, ap, mp = O.rasterization(*args, sh_degree=2, packed=True, rasterize_mode="antialiased")
    assert torch.equal(rd, rp) and torch.equal(ad, ap)
    assert torch.equal(mp["gaussian_ids"], (md["radii"].reshape(-1) > 0).nonzero()[:, 0] % 1500)


def test_oracle_backward_matches_finite_differences_fp64():
    s = make_scene(12, 1, 1, 32, 32, seed=7)
    for k in ("means", "quats", "scales", "opacities", "colors"):
        s[k] = s[k].double()
    s["scales"] = s["scales"] * 0.5
    g = torch.Generator().manual_seed(1)
    v = torch.rand(1, 32, 32, 3, generator=g, dtype=torch.float64)

    def loss(means, quats, scales, opac, cols):
        r, a, _ = O.rasterization(means, quats, scales, opac, cols, s["viewmats"], s["Ks"], 32, 32, sh_degree=1,
                                  packed=False, rasterize_mode="antialiased")
        return (r * v).sum() + 0.3 * a.sum()

    inputs = [s[k].clone().requires_grad_() for k in ("means", "quats", "scales", "opacities", "colors")]
    grads = torch.autograd.grad(loss(*inputs), inputs)
    eps = 1e-6
    for t, gr in zip(inputs, grads):
        flat = t.detach().reshape(-1)
        for idx in range(0, flat.numel(), max(1, flat.numel() // 7)):
            def f(delta):
                tt = [x.detach().clone() for x in inputs]
                which = [i for i, x in enumerate(inputs) if x is t][0]
                tt[which].reshape(-1)[idx] += delta
                return loss(*tt).item()
            fd = (f(eps) - f(-eps)) / (2 * eps)
            assert abs(fd - gr.reshape(-1)[idx].item()) < 1e-5 * max(1.0, abs(fd)), (idx, fd, gr.reshape(-1)[idx].item())


def test_oracle_reproduces_committed_golden():
    path = os.path.join(HERE, "golden", "scene_small.npz")
    gold = np.load(path)
    from tests.golden.make_golden import compute, SPEC
    out = compute(SPEC)
    for k in gold.files:
        a, b = gold[k], out[k]
        if a.dtype.kind in "iu":
            assert np.array_equal(a, b), k
        else:
            assert np.allclose(a, b, rtol=1e-5, atol=1e-6), k
