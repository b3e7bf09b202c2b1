"""CPU tests of the trainer shell's host logic: command line of the reference's Train-Stage1
component, COLMAP model I/O, and the checkpoint -> PLY export against a numpy restatement of the reference's
gsplat_pt_to_ply.py (source/container/src/pipeline/post_processing/gsplat_pt_to_ply.py:52-75)."""
import os
import struct

import numpy as np
import pytest
import torch

from pipeline_pointcloud_b200 import colmap, export
from pipeline_pointcloud_b200.simple_trainer import parse_args


def test_reference_command_line_is_accepted():
    # the exact argument list source/container/src/main.py:1328-1338 builds (4 GPUs, MAX_STEPS 30000)
    a = parse_args(["default", "--max_steps", "30000", "--result-dir", "/tmp/out", "--data_factor", "1", "--steps_scaler", "0.25",
                    "--disable_viewer", "--packed", "--batch-size", "1", "--data-dir", "/data/set"])
    assert a.strategy == "default" and a.packed and a.disable_viewer and a.batch_size == 1 and a.data_factor == 1
    assert a.max_steps == 7500 and a.save_steps == [1750, 7500] and a.sh_degree_interval == 250
    assert a.result_dir == "/tmp/out" and a.data_dir == "/data/set"
    m = parse_args(["mcmc", "--data-dir", "x"])              # gsplat's mcmc preset
    assert (m.init_opa, m.init_scale, m.opacity_reg, m.scale_reg, m.cap_max) == (0.5, 0.1, 0.01, 0.01, 1_000_000)
    assert (a.init_opa, a.init_scale, a.opacity_reg, a.scale_reg) == (0.1, 1.0, 0.0, 0.0)
    assert parse_args(["default", "--data-dir", "x", "--batch-size", "2"]).batch_size == 2
    with pytest.raises(ValueError):
        parse_args(["default", "--data-dir", "x", "--batch-size", "0"])
    with pytest.raises(SystemExit):
        parse_args(["default"])                      # --data-dir is required


def test_colmap_model_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    n_img = 5
    Ks = np.stack([np.array([[500.0 + i, 0, 160], [0, 510.0 + i, 120], [0, 0, 1]]) for i in range(n_img)])
    views = []
    for i in range(n_img):
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] *= -1
        V = np.eye(4)
        V[:3, :3], V[:3, 3] = Q, rng.normal(size=3)
        views.append(V)
    names = [f"img_{i:03d}.png" for i in range(n_img)]
    xyz, rgb = rng.normal(size=(40, 3)), rng.integers(0, 256, size=(40, 3), dtype=np.uint8)
    sparse = str(tmp_path / "sparse" / "0")
    colmap.write_model(sparse, Ks, np.stack(views), names, 320, 240, xyz, rgb)
    assert colmap.find_sparse_dir(str(tmp_path)) == sparse
    cams, imgs = colmap.read_cameras(sparse + "/cameras.bin"), colmap.read_images(sparse + "/images.bin")
    p, c = colmap.read_points3d(sparse + "/points3D.bin")
    assert np.allclose(p, xyz) and np.array_equal(c, rgb)
    assert len(cams) == n_img and len(imgs) == n_img
    for i in range(n_img):
        im = imgs[i + 1]
        assert im.name == names[i] and cams[im.camera_id].model == "PINHOLE"
        assert np.allclose(im.world_to_cam(), views[i], atol=1e-9)
        assert np.allclose(cams[im.camera_id].K(), Ks[i]) and cams[im.camera_id].distortion().size == 0
    # a camera model with distortion parameters is parsed (the trainer then refuses it)
    with open(sparse + "/cameras.bin", "wb") as f:
        f.write(struct.pack("<Q", 1) + struct.pack("<iiQQ", 1, 2, 320, 240) + struct.pack("<dddd", 400.0, 160.0, 120.0, 0.05))
    cam = colmap.read_cameras(sparse + "/cameras.bin")[1]
    assert cam.model == "SIMPLE_RADIAL" and np.allclose(cam.distortion(), [0.05]) and cam.K()[0, 0] == 400.0


def reference_ply_channels(splats):
    """numpy restatement of gsplat_pt_to_ply.py:52-73 (column order and SH transposition)."""
    pos = splats["means"].numpy()
    cols = [("x", pos[:, 0]), ("y", pos[:, 1]), ("z", pos[:, 2])] + [(k, np.zeros(len(pos), np.float32)) for k in ("nx", "ny", "nz")]
    f_dc = np.transpose(splats["sh0"].numpy(), (0, 2, 1)).reshape(len(pos), -1)
    cols += [(f"f_dc_{i}", f_dc[:, i]) for i in range(f_dc.shape[1])]
    f_rest = np.transpose(splats["shN"].numpy(), (0, 2, 1)).reshape(len(pos), -1)
    cols += [(f"f_rest_{i}", f_rest[:, i]) for i in range(f_rest.shape[1])]
    cols.append(("opacity", splats["opacities"].numpy()))
    cols += [(f"scale_{i}", splats["scales"].numpy()[:, i]) for i in range(3)]
    cols += [(f"rot_{i}", splats["quats"].numpy()[:, i]) for i in range(4)]
    return cols


def test_export_ply_matches_reference_layout(tmp_path):
    g = torch.Generator().manual_seed(0)
    N = 37
    splats = {"means": torch.randn(N, 3, generator=g), "scales": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
              "opacities": torch.randn(N, generator=g), "sh0": torch.randn(N, 1, 3, generator=g), "shN": torch.randn(N, 15, 3, generator=g)}
    ck = tmp_path / "ckpts"
    ck.mkdir()
    torch.save({"step": 9, "splats": {k: torch.zeros_like(v) for k, v in splats.items()}}, ck / "ckpt_1749_rank0.pt")
    torch.save({"step": 99, "splats": splats}, ck / "ckpt_7499_rank0.pt")        # lexicographically last -> the one exported
    out = str(tmp_path / "splat.ply")
    assert export.main([str(ck), out]) == 0
    raw = open(out, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    ref = reference_ply_channels(splats)
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {N}"]
    assert lines[3:] == [f"property float {k}" for k, _ in ref]
    assert len(ref) == 3 + 3 + 3 + 45 + 1 + 3 + 4
    table = np.stack([np.asarray(v, np.float32).reshape(-1) for _, v in ref], axis=1)      # one interleaved row per Gaussian
    assert body == table.astype("<f4").tobytes()
    back = export.read_ply(out)
    assert list(back) == [k for k, _ in ref] and np.array_equal(back["f_rest_17"], table[:, 9 + 17])
    # f_rest is channel-major: f_rest_0..14 = red coefficients 1..15
    assert np.array_equal(back["f_rest_3"], splats["shN"][:, 3, 0].numpy()) and np.array_equal(back["f_rest_15"], splats["shN"][:, 0, 1].numpy())


def test_mcmc_relocation_preserves_opacity_and_closed_form():
    import math
    from pipeline_pointcloud_b200.simple_trainer import compute_relocation
    o = torch.tensor([0.9, 0.5, 0.2, 0.7])
    sc = torch.tensor([[1.0, 2.0, 3.0]] * 4)
    N = torch.tensor([1, 2, 3, 80])                                   # 80 is capped at 51
    new_o, new_s = compute_relocation(o, sc, N)
    n_eff = torch.tensor([1.0, 2.0, 3.0, 51.0])
    assert torch.allclose(1 - (1 - new_o) ** n_eff, o, atol=1e-6)   # N copies composite to the original opacity
    assert torch.allclose(new_o[0], o[0]) and torch.allclose(new_s[0], sc[0])          # N = 1: unchanged
    x = new_o[1].item()                                               # N = 2: denominator x + (x - x^2 / sqrt 2)
    assert abs(new_s[1, 0].item() - 0.5 / (x + x - x * x / math.sqrt(2))) < 1e-6
    assert (new_s[1:] < sc[1:]).all() and torch.isfinite(new_s).all()


def _toy_splats(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    return {"means": torch.randn(n, 3, generator=g).requires_grad_(), "scales": (torch.randn(n, 3, generator=g) * 0.3 - 6.0).requires_grad_(),
            "quats": torch.randn(n, 4, generator=g).requires_grad_(), "opacities": torch.randn(n, generator=g).requires_grad_(),
            "sh0": torch.randn(n, 1, 3, generator=g).requires_grad_(), "shN": torch.zeros(n, 15, 3).requires_grad_()}


def test_default_strategy_clone_split_prune_and_optimizer_state():
    """Densification bookkeeping on CPU tensors (the strategy is plain torch): clones keep their parameters, splits
    produce two shrunken children, pruned rows vanish, and the Adam moments follow the rows (zeros for the new ones)."""
    from pipeline_pointcloud_b200.simple_trainer import DefaultStrategy
    from pipeline_pointcloud_b200.training import FusedAdam
    n = 50
    sp = _toy_splats(n)
    with torch.no_grad():
        sp["scales"][10:20] = -2.0                     # exp(-2) = 0.14 > 0.01 * scene_scale -> split candidates
        sp["opacities"][40:45] = -9.0                  # sigmoid < 0.005 -> pruned
    opt = FusedAdam([{"params": [sp[k]], "lr": 1e-3, "name": k} for k in sp])
    for k, p in sp.items():                            # pretend some steps happened: recognisable moments
        opt.state[p] = {"step": 7, "exp_avg": torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape) + 1,
                        "exp_avg_sq": torch.ones_like(p) * 2}
    st = DefaultStrategy(scene_scale=1.0, refine_start=0, refine_every=1)
    st.grad2d, st.count = torch.zeros(n), torch.ones(n)
    st.grad2d[0:5] = 1.0                               # small + high gradient -> duplicated
    st.grad2d[10:20] = 1.0                             # large + high gradient -> split
    old = {k: v.detach().clone() for k, v in sp.items()}
    stats = st.refine(sp, opt, step=1)
    assert stats["duplicated"] == 5 and stats["split"] == 10 and stats["pruned"] == 5
    N = sp["means"].shape[0]
    assert N == stats["n"] == n - 10 + 5 + 20 - 5
    for k, p in sp.items():
        assert p.shape[0] == N and p.requires_grad and p.is_leaf
        s_ = opt.state[p]
        assert s_["step"] == 7 and s_["exp_avg"].shape == p.shape and s_["exp_avg_sq"].shape == p.shape
        assert any(q is p for g_ in opt.param_groups for q in g_["params"])
    # survivors first (rows 0..9, 20..39, 45..49), then the 5 clones, then 2 x 10 children
    keep = [i for i in range(n) if not (10 <= i < 20 or 40 <= i < 45)]
    assert torch.equal(sp["quats"][: len(keep)], old["quats"][keep])
    assert torch.equal(sp["means"][len(keep): len(keep) + 5], old["means"][0:5])
    kids = sp["scales"][len(keep) + 5:]
    assert kids.shape[0] == 20 and torch.allclose(torch.exp(kids[:10]), torch.exp(old["scales"][10:20]) / 1.6)
    assert opt.state[sp["means"]]["exp_avg"][len(keep):].abs().sum() == 0          # new rows start with zero moments
    assert opt.state[sp["opacities"]]["exp_avg"][: len(keep)].min() > 0          # survivors keep theirs
    # opacity reset: capped at logit(2 * prune_opa), moments zeroed
    with torch.no_grad():
        sp["opacities"].fill_(3.0)
    st2 = DefaultStrategy(scene_scale=1.0, reset_every=4)
    st2.grad2d = torch.zeros(N)
    st2.count = torch.zeros(N)
    stats = st2.refine(sp, opt, step=8)
    assert stats.get("opacity_reset") == 1
    assert torch.allclose(torch.sigmoid(sp["opacities"]), torch.full((N,), 0.01), atol=1e-6)
    assert opt.state[sp["opacities"]]["exp_avg"].abs().sum() == 0


# ----------------------------------------------------------------------------------------------------------
# launch boundary: the reference starts plain `python gsplat/examples/simple_trainer.py` (no torchrun)
# ----------------------------------------------------------------------------------------------------------
REF_ARGS = ["default", "--max_steps", "30000", "--result-dir", "OUT", "--data_factor", "1", "--steps_scaler", "0.25",
            "--disable_viewer", "--packed", "--batch-size", "1", "--data-dir", "DATA"]      # main.py:1328-1338, 4 GPUs


def test_launch_plan_follows_upstream_cli():
    from pipeline_pointcloud_b200.simple_trainer import launch_plan
    # the environment source/container/src/main.py:645-653 leaves behind on a 4-GPU instance: WORLD_SIZE counts nodes
    ref_env = {"CUDA_VISIBLE_DEVICES": "0,1,2,3", "MASTER_ADDR": "localhost", "MASTER_PORT": "29500", "WORLD_SIZE": "1",
               "RANK": "0", "LOCAL_RANK": "0"}
    p = launch_plan(ref_env, n_local_gpus=4)
    assert p["mode"] == "spawn" and p["world"] == 4 and p["n_local"] == 4 and p["node_rank"] == 0
    assert p["master_addr"] == "127.0.0.1" and p["master_port"] == 29500
    # multi-container SageMaker (main.py:628-642): one GPU per host, WORLD_SIZE = hosts, RANK = host index
    p = launch_plan({"WORLD_SIZE": "3", "RANK": "2", "LOCAL_RANK": "0", "MASTER_ADDR": "algo-1", "MASTER_PORT": "29500"}, 1)
    assert p["mode"] == "spawn" and p["world"] == 3 and p["node_rank"] == 2 and p["n_local"] == 1 and p["master_addr"] == "algo-1"
    # single GPU, nothing set
    assert launch_plan({}, 1)["mode"] == "single" and launch_plan({}, 0)["mode"] == "single"
    # under torchrun the process is a worker already
    p = launch_plan({"TORCHELASTIC_RUN_ID": "x", "RANK": "5", "WORLD_SIZE": "8", "LOCAL_RANK": "5", "LOCAL_WORLD_SIZE": "8"}, 8)
    assert p == {"mode": "worker", "rank": 5, "world": 8, "local_rank": 5}


def test_optimizer_scaling_and_view_order():
    from pipeline_pointcloud_b200.simple_trainer import optimizer_hparams, view_of_step
    h1, h4 = optimizer_hparams(1), optimizer_hparams(4)
    assert h1 == {"lr_scale": 1.0, "eps": 1e-15, "betas": (0.9, 0.999)}
    assert h4["lr_scale"] == 2.0 and h4["eps"] == 0.5e-15
    assert abs(h4["betas"][0] - 0.6) < 1e-12 and abs(h4["betas"][1] - 0.996) < 1e-12
    order = np.random.default_rng(0).permutation(10)
    for step in range(5):            # the ranks of one step render different views; two steps cover the epoch once
        views = [view_of_step(order, step, r, 4) for r in range(4)]
        assert len(set(views)) == 4
    assert sorted(view_of_step(order, s, r, 2) for s in range(5) for r in range(2)) == list(range(10))
    # --batch-size 2 on 2 ranks: four different views per step; an epoch of 12 views takes three steps
    order = np.random.default_rng(0).permutation(12)
    per_step = [[view_of_step(order, s, r, 2, 2, b) for r in range(2) for b in range(2)] for s in range(3)]
    assert all(len(set(v)) == 4 for v in per_step) and sorted(x for v in per_step for x in v) == list(range(12))
    assert optimizer_hparams(2, 2) == optimizer_hparams(4)          # BS = batch size x world


def _small_colmap(root):
    rng = np.random.default_rng(1)
    n_img, W, H = 3, 64, 48
    import cv2
    os.makedirs(os.path.join(root, "images"), exist_ok=True)
    names = []
    for i in range(n_img):
        names.append(f"v{i}.png")
        cv2.imwrite(os.path.join(root, "images", names[-1]), rng.integers(0, 255, size=(H, W, 3), dtype=np.uint8))
    Ks = np.stack([np.array([[50.0, 0, W / 2], [0, 50.0, H / 2], [0, 0, 1]])] * n_img)
    V = np.stack([np.eye(4)] * n_img)
    V[:, 2, 3] = 4.0
    colmap.write_model(os.path.join(root, "sparse", "0"), Ks, V, names, W, H, rng.normal(size=(30, 3)),
                       rng.integers(0, 255, size=(30, 3), dtype=np.uint8))


REF_SRC = "/root/reference/source/container/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree is only present in the build container")
def test_train_stage1_component_through_the_reference_runner(tmp_path, monkeypatch, capfd):
    """The reference's own operator API for this path (source/container/src/pipeline/pipeline.py:152-235): the unchanged
    Train-Stage1 component -- command "gsplat/examples/simple_trainer.py", the argument list of main.py:1328-1338, cwd = the
    directory that holds gsplat/ (here: shim/) -- run by the reference's Pipeline.run_component.  No GPU in this container:
    the dry-run environment switch exercises argv / env handling and the exit-code contract; the real run is
    tests/test_gpu_trainer.py::test_reference_launch_spawns_one_worker_per_gpu."""
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("ref_pipeline", os.path.join(REF_SRC, "pipeline", "pipeline.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    data, out = str(tmp_path / "data"), str(tmp_path / "out")
    _small_colmap(data)
    monkeypatch.chdir(tmp_path)                      # Pipeline() opens "<name>-pipeline-log" in the cwd
    for k in ("TORCHELASTIC_RUN_ID", "LOCAL_WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("WORLD_SIZE", "1"); monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setenv("MASTER_ADDR", "localhost"); monkeypatch.setenv("MASTER_PORT", "29500")
    pl = ref.Pipeline(name="b2s-test", uuid="u", num_threads="2", num_gpus="2", log_verbosity="error")
    args = [a.replace("OUT", out).replace("DATA", data) for a in REF_ARGS]
    args[args.index("--steps_scaler") + 1] = "0.5"
    pl.create_component(name="Train-Stage1", comp_type=ref.ComponentType.transform, comp_environ=ref.ComponentEnvironment.python,
                        command="gsplat/examples/simple_trainer.py", args=args + ["CUDA_VISIBLE_DEVICES=0,1"],
                        cwd=os.path.join(root, "shim"), requires_gpu=True)
    # 1. success contract: exit code 0 -> run_component returns
    monkeypatch.setenv("B2S_TRAINER_DRY_RUN", "1")
    pl.run_component(0)
    seen = capfd.readouterr().out
    assert "dry run: CUDA_VISIBLE_DEVICES=0,1" in seen and "3 views, 30 SfM points, max_steps 15000" in seen
    # 2. failure contract: the trainer raises (no CUDA device here) -> non-zero exit -> the runner stops the pipeline
    monkeypatch.delenv("B2S_TRAINER_DRY_RUN")
    if not torch.cuda.is_available():
        with pytest.raises(SystemExit) as e:
            pl.run_component(0)
        assert e.value.code == 1
    # 3. argv sanitising happens before anything is launched
    pl.create_component(name="bad", comp_type=ref.ComponentType.transform, comp_environ=ref.ComponentEnvironment.python,
                        command="gsplat/examples/simple_trainer.py", args=["default", "--data-dir", "$(rm -rf x)"],
                        cwd=os.path.join(root, "shim"), requires_gpu=True)
    with pytest.raises(ValueError):
        pl.run_component(1)


def test_shim_shadows_gsplat(monkeypatch):
    """`from gsplat.rendering import rasterization` (nerfstudio's import, behind main.py:1308-1316) and the stage operators
    resolve to this package when shim/ is on the module path."""
    import importlib
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "shim"))
    for m in [m for m in sys.modules if m == "gsplat" or m.startswith("gsplat.")]:
        monkeypatch.delitem(sys.modules, m)
    import pipeline_pointcloud_b200 as b2s
    gs = importlib.import_module("gsplat")
    rend = importlib.import_module("gsplat.rendering")
    assert rend.rasterization is b2s.rasterization and gs.rasterization is b2s.rasterization
    for name in ("fully_fused_projection", "spherical_harmonics", "isect_tiles", "isect_offset_encode", "rasterize_to_pixels",
                 "quat_scale_to_covar_preci"):
        assert getattr(gs, name) is getattr(b2s, name)
    assert os.path.isfile(os.path.join(root, "shim", "gsplat", "examples", "simple_trainer.py"))
    with pytest.raises(Exception):                                   # no CPU path behind the shim either
        t = torch.zeros(4, 3)
        rend.rasterization(t, torch.zeros(4, 4), t, torch.zeros(4), t, torch.eye(4)[None], torch.eye(3)[None], 16, 16)
