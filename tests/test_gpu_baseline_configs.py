"""GPU parity at the BASELINE.json configurations themselves (no parity test ran at any BASELINE config).

* config 1 (100 k Gaussians, SH degree 0, 256x256, classic, forward): the whole frame against the CPU oracle;
* config 3 (5.8 M Gaussians, SH degree 3, 1297x840, antialiased, fwd+bwd): the CUDA path renders and differentiates the
  whole frame; the oracle does a central sub-frustum of it with whole-frame geometry (oracle/window_check.py), so the
  window's slice of the sorted key list is compared bit for bit, its pixels at 1e-4 and the gradients at 1e-3;
* which side is closer to the fp64 oracle, and what the error on the knife-edge pixels is.
Every test appends its figures to gpurun_out/parity_report.jsonl (kept under profiles/)."""
import math

import pytest
import torch

from oracle import torch_oracle as O
from oracle import window_check as WC
from tests.util import LEAVES, call, grad_report, report, scene_args

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b2s():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pipeline_pointcloud_b200 as pkg
    return pkg


def test_config1_whole_frame_against_oracle(b2s):
    N, deg, W, H = 100_000, 0, 256, 256
    a = scene_args(N, deg, 1, W, H, seed=0)
    with torch.no_grad():
        ref_r, ref_a, ref_m = call(O.rasterization, a, W, H, sh_degree=deg, packed=False, rasterize_mode="classic", return_fragile=True)
        g = {k: v.cuda() for k, v in a.items()}
        for packed in (False, True):
            r, al, m = call(b2s.rasterization, g, W, H, sh_degree=deg, packed=packed, rasterize_mode="classic")
            flat = m["flatten_ids"] if not packed else m["gaussian_ids"][m["flatten_ids"].long()]
            assert torch.equal(m["isect_ids"].cpu(), ref_m["isect_ids"])
            assert torch.equal(flat.cpu().long(), ref_m["flatten_ids"].long())
            assert torch.equal(m["isect_offsets"].cpu(), ref_m["isect_offsets"])
            fr = ref_m["fragile"]
            dr = (r.cpu() - ref_r).abs().amax(-1)
            da = (al.cpu() - ref_a).abs().amax(-1)
            rep = {"packed": packed, "n_isects": int(ref_m["isect_ids"].numel()), "fragile_frac": float(fr.float().mean()),
                   "render_max_err": float(dr[~fr].max()), "alpha_max_err": float(da[~fr].max()),
                   "render_max_err_on_fragile": float(dr[fr].max()) if fr.any() else 0.0,
                   "alpha_max_err_on_fragile": float(da[fr].max()) if fr.any() else 0.0}
            report("config1_whole_frame", rep)
            assert rep["render_max_err"] <= 1e-4 and rep["alpha_max_err"] <= 1e-4, rep
            assert rep["fragile_frac"] < 0.05 and rep["render_max_err_on_fragile"] < 2e-2, rep


@pytest.mark.parametrize("workload", [(5_800_000, 3, 1297, 840, "antialiased"), (1_000_000, 3, 800, 800, "classic")],
                         ids=["config3", "config2"])
def test_full_size_subfrustum_against_oracle(b2s, workload):
    N, deg, W, H, mode = workload
    a = scene_args(N, deg, 1, W, H, seed=0)
    window = WC.central_window(W, H)
    sub, idx = WC.select_window_inputs(a, W, H, window)
    v_r, v_a = WC.window_cotangents(W, H, window, seed=1)
    ref = WC.oracle_window_step(sub, W, H, deg, mode, window, v_r, v_a, zero_fragile=True)
    assert ref["n_isects"] > 50_000
    for packed in (False, True):
        g = {k: (v.cuda().requires_grad_() if k in LEAVES else v.cuda()) for k, v in a.items()}
        r, al, m = call(b2s.rasterization, g, W, H, sh_degree=deg, packed=packed, rasterize_mode=mode)
        torch.autograd.backward([r, al], [ref["v_render"].cuda(), ref["v_alpha"].cuda()])
        flat = m["flatten_ids"] if not packed else m["gaussian_ids"][m["flatten_ids"].long()]
        rep = WC.compare_window(ref, idx, r.detach().cpu(), al.detach().cpu(), m["isect_ids"].cpu(), flat.cpu(),
                                {k: g[k].grad.cpu() for k in LEAVES}, W, H, window)
        rep["packed"] = packed
        report(f"subfrustum_N{N}", rep)
        assert rep["keys_bit_equal"] and rep["ids_bit_equal"], rep          # integer work: bit-exact
        assert rep["render_max_err"] <= 1e-4 and rep["alpha_max_err"] <= 1e-4, rep
        assert max(rep["grad_rel_err"].values()) < 1e-3, rep
        assert rep["grad_outside_selection_max"] == 0.0, rep               # nothing outside the oracle's selection was touched
        assert rep["fragile_frac"] < 0.05 and rep["render_max_err_on_fragile"] < 2e-2, rep
        del g, r, al, m


def test_which_side_is_closer_to_fp64(b2s):
    """CUDA fp32 and oracle fp32 against the oracle in fp64 on the same inputs: the kernels must not be the worse fp32
    implementation by more than a small factor, and the error ON the knife-edge pixels is reported (they are excluded from
    the 1e-4 bar, not from scrutiny)."""
    N, deg, C, W, H = 4000, 3, 2, 112, 80
    a = scene_args(N, deg, C, W, H, seed=11)
    kw = dict(sh_degree=deg, packed=False, rasterize_mode="antialiased")
    a64 = {k: v.double() for k, v in a.items()}
    gen = torch.Generator().manual_seed(3)
    v_r, v_a = torch.rand(C, H, W, 3, generator=gen), torch.rand(C, H, W, 1, generator=gen)

    def run(fn, args, dt):
        leaves = {k: args[k].detach().clone().requires_grad_() for k in LEAVES}
        full = dict(args); full.update(leaves)
        r, al, m = call(fn, full, W, H, return_fragile=True, **kw) if fn is O.rasterization else call(fn, full, W, H, **kw)
        return r, al, m, leaves

    r64, a64_, m64, l64 = run(O.rasterization, a64, torch.float64)
    r32, a32_, m32, l32 = run(O.rasterization, a, torch.float32)
    g = {k: v.cuda() for k, v in a.items()}
    rg, ag, mg, lg = run(b2s.rasterization, g, torch.float32)
    fragile = m64["fragile"] | m32["fragile"]
    keep = (~fragile)[..., None]
    for (r, al, dt) in ((r64, a64_, torch.float64), (r32, a32_, torch.float32)):
        torch.autograd.backward([r, al], [(v_r * keep).to(dt), (v_a * keep).to(dt)])
    torch.autograd.backward([rg, ag], [(v_r * keep).cuda(), (v_a * keep).cuda()])
    ok = ~fragile
    rep = {"fragile_frac": float(fragile.float().mean())}
    for name, x, y in (("render", rg.detach().cpu(), r32.detach()), ("alpha", ag.detach().cpu(), a32_.detach())):
        ref = (r64 if name == "render" else a64_).detach()
        e_gpu, e_o32 = (x.double() - ref).abs().amax(-1), (y.double() - ref).abs().amax(-1)
        rep[name] = {"cuda_vs_fp64": float(e_gpu[ok].max()), "oracle32_vs_fp64": float(e_o32[ok].max()),
                     "cuda_vs_fp64_on_fragile": float(e_gpu[fragile].max()) if fragile.any() else 0.0,
                     "oracle32_vs_fp64_on_fragile": float(e_o32[fragile].max()) if fragile.any() else 0.0}
        assert rep[name]["cuda_vs_fp64"] <= 1e-4, rep
    for k in LEAVES:
        ref = l64[k].grad
        e_gpu = float((lg[k].grad.cpu().double() - ref).norm() / ref.norm())
        e_o32 = float((l32[k].grad.double() - ref).norm() / ref.norm())
        rep["grad_" + k] = {"cuda_vs_fp64": e_gpu, "oracle32_vs_fp64": e_o32, "cuda_is_closer": e_gpu <= e_o32}
        assert e_gpu < 1e-3, (k, rep)
        assert e_gpu < max(10.0 * e_o32, 2e-4), (k, rep)        # never the clearly worse fp32 implementation
    report("fp64_which_side_is_closer", rep)


def test_quat_scale_to_covar_preci(b2s):
    """The stand-alone op (the training path fuses it into the projection)."""
    torch.manual_seed(0)
    q, s = torch.randn(5000, 4), torch.rand(5000, 3) * 2 + 0.05
    cov_ref, pre_ref = O.quat_scale_to_covar_preci(q.double(), s.double())
    cov, pre = b2s.quat_scale_to_covar_preci(q.cuda(), s.cuda())
    assert cov.shape == (5000, 3, 3) and pre.shape == (5000, 3, 3)
    assert ((cov.cpu().double() - cov_ref).abs() / cov_ref.abs().clamp_min(1.0)).max() < 1e-5
    assert ((pre.cpu().double() - pre_ref).abs() / pre_ref.abs().clamp_min(1.0)).max() < 1e-4
    eye = torch.bmm(cov.cpu().double(), pre.cpu().double())
    assert (eye - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-3
    tri_c, tri_p = b2s.quat_scale_to_covar_preci(q.cuda(), s.cuda(), triu=True)
    assert tri_c.shape == (5000, 6) and torch.equal(tri_c[:, 1], cov.reshape(-1, 9)[:, 1])
    only_c, none_p = b2s.quat_scale_to_covar_preci(q.cuda(), s.cuda(), compute_preci=False)
    assert none_p is None and torch.equal(only_c, cov)
