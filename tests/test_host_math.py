"""CPU checks of the kernels' per-element math header (csrc/b2s_math.cuh compiled with g++) against the
oracle: forward projection + tile rectangles bit-exact, hand-derived VJPs against oracle autograd."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import torch_oracle as O
from pipeline_pointcloud_b200.scene import make_scene
from tests.host_math.build import build

F = ctypes.POINTER(ctypes.c_float)
I = ctypes.POINTER(ctypes.c_int)


def fp(t):
    return ctypes.cast(t.data_ptr(), F)


def ip(t):
    return ctypes.cast(t.data_ptr(), I)


@pytest.fixture(scope="module")
def hm():
    return ctypes.CDLL(build())


def scene(N=5000, C=3, W=200, H=120, seed=0, deg=3):
    s = make_scene(N, deg, C, W, H, seed=seed)
    # push some Gaussians behind / beside the cameras and make some huge so every cull branch is hit
    s["means"][:50] *= 4.0
    s["scales"][50:80] *= 30.0
    return s, W, H


@pytest.mark.parametrize("seed", [0, 1])
def test_projection_forward_bit_exact(hm, seed):
    s, W, H = scene(seed=seed)
    N, C = s["means"].shape[0], s["viewmats"].shape[0]
    radii = torch.zeros(C, N, dtype=torch.int32)
    means2d = torch.zeros(C, N, 2)
    depths = torch.zeros(C, N)
    conics = torch.zeros(C, N, 3)
    comps = torch.zeros(C, N)
    rects = torch.zeros(C, N, 4, dtype=torch.int32)
    tw, th = (W + 15) // 16, (H + 15) // 16
    hm.hm_project_fwd(ctypes.c_long(N), C, fp(s["means"]), fp(s["quats"]), fp(s["scales"]), fp(s["viewmats"]), fp(s["Ks"]),
                      W, H, ctypes.c_float(0.3), ctypes.c_float(0.01), ctypes.c_float(1e10), ctypes.c_float(0.0),
                      16, tw, th, ip(radii), fp(means2d), fp(depths), fp(conics), fp(comps), ip(rects))
    o_radii, o_means2d, o_depths, o_conics, o_comps = O.fully_fused_projection(
        s["means"], s["quats"], s["scales"], s["viewmats"], s["Ks"], W, H, calc_compensations=True)
    assert (o_radii > 0).sum() > 1000 and (o_radii == 0).sum() > 20
    assert torch.equal(radii, o_radii)
    # bit-exact: compare the raw bit patterns
    for a, b in [(means2d, o_means2d), (depths, o_depths), (conics, o_conics), (comps, o_comps)]:
        assert torch.equal(a.view(torch.int32), b.contiguous().view(torch.int32))
    tmin, tmax = O.tile_rects(o_means2d.reshape(-1, 2), o_radii.reshape(-1), 16, tw, th)
    vis = (o_radii > 0).reshape(-1)
    r = rects.reshape(-1, 4)
    assert torch.equal(r[vis][:, :2], tmin[vis]) and torch.equal(r[vis][:, 2:], tmax[vis])


def test_projection_vjp_matches_autograd(hm):
    s, W, H = scene(N=3000, seed=2)
    N, C = s["means"].shape[0], s["viewmats"].shape[0]
    g = torch.Generator().manual_seed(5)
    means = s["means"].double().requires_grad_()
    quats = s["quats"].double().requires_grad_()
    scales = s["scales"].double().requires_grad_()
    radii, means2d, depths, conics, comps = O.fully_fused_projection(
        means, quats, scales, s["viewmats"], s["Ks"], W, H, calc_compensations=True)
    v_means2d = torch.randn(C, N, 2, generator=g)
    v_depths = torch.randn(C, N, generator=g)
    v_conics = torch.randn(C, N, 3, generator=g)
    v_comps = torch.randn(C, N, generator=g)
    loss = (means2d * v_means2d).sum() + (depths * v_depths).sum() + (conics * v_conics).sum() + (comps * v_comps).sum()
    loss.backward()
    vm, vq, vs = torch.zeros(N, 3), torch.zeros(N, 4), torch.zeros(N, 3)
    hm.hm_project_bwd(ctypes.c_long(N), C, fp(s["means"]), fp(s["quats"]), fp(s["scales"]), fp(s["viewmats"]), fp(s["Ks"]),
                      W, H, ctypes.c_float(0.3), ctypes.c_float(0.01), ctypes.c_float(1e10), ctypes.c_float(0.0),
                      fp(v_means2d), fp(v_depths), fp(v_conics), fp(v_comps), fp(vm), fp(vq), fp(vs))
    for name, got, ref in [("means", vm, means.grad), ("quats", vq, quats.grad), ("scales", vs, scales.grad)]:
        ref = ref.float()
        err = (got - ref).abs()
        tol = 1e-3 * ref.abs() + 1e-4 * ref.abs().mean()
        frac_bad = (err > tol).float().mean().item()
        rel = (got - ref).norm() / ref.norm()
        assert rel < 1e-4, (name, rel)
        assert frac_bad < 2e-3, (name, frac_bad)


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_forward_and_vjp(hm, deg):
    g = torch.Generator().manual_seed(deg)
    M, Kt = 4000, 16
    dirs = torch.randn(M, 3, generator=g) * 3.0
    coeffs = torch.randn(M, Kt, 3, generator=g)
    out = torch.zeros(M, 3)
    hm.hm_sh_fwd(ctypes.c_long(M), deg, Kt, fp(dirs), fp(coeffs), fp(out))
    d64 = dirs.double().requires_grad_()
    c64 = coeffs.double().requires_grad_()
    ref = O.spherical_harmonics(deg, d64, c64)
    assert (out - ref.float()).abs().max() < 2e-5
    v_out = torch.randn(M, 3, generator=g)
    (ref * v_out).sum().backward()
    vc, vd = torch.zeros(M, Kt, 3), torch.zeros(M, 3)
    hm.hm_sh_bwd(ctypes.c_long(M), deg, Kt, fp(dirs), fp(coeffs), fp(v_out), fp(vc), fp(vd))
    assert (vc - c64.grad.float()).abs().max() < 2e-5
    if deg == 0:
        assert vd.abs().max() == 0          # colour does not depend on the view direction
    else:
        assert (vd - d64.grad.float()).norm() / d64.grad.norm() < 1e-4
