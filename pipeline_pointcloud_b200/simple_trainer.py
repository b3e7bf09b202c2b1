"""Trainer shell around the rasterization path: a drop-in for the command line the reference's
Train-Stage1 component runs,

    python gsplat/examples/simple_trainer.py {default|mcmc} --max_steps N --result-dir DIR --data_factor 1
        --steps_scaler S --disable_viewer --packed --batch-size 1 --data-dir DATASET

(source/container/src/main.py:1328-1347), producing the checkpoint the export stage reads:
``DIR/ckpts/ckpt_{step}_rank{r}.pt`` = ``{"step", "splats": {means [N,3], scales [N,3] (log), quats [N,4] (wxyz),
opacities [N] (logit), sh0 [N,1,3], shN [N,K-1,3]}}`` (source/container/src/pipeline/post_processing/gsplat_pt_to_ply.py:38-73).

Every iteration runs on this package's kernels: ``rasterization()`` forward + backward, the fused L1+SSIM loss and the
fused (Selective)Adam update.  Densification follows gsplat's ``DefaultStrategy`` (clone / split / prune / opacity
reset) and ``MCMCStrategy`` (relocation, 5 % growth, position noise) with torch tensor ops -- they run every 100th step
and are not on the per-iteration hot path.

Multi-GPU follows upstream's ``gsplat.distributed.cli``: the reference launches plain ``python simple_trainer.py`` with
``CUDA_VISIBLE_DEVICES=0..n-1``, ``WORLD_SIZE=1`` (the number of *nodes*), ``RANK=0`` (the node rank) and
``--steps_scaler 1/n`` (source/container/src/main.py:620-655, :1321-1347), so ``main()`` itself spawns one worker per
visible GPU (``launch``); under torchrun the ranks torchrun made are used as they are.  Gaussians are sharded over the ranks
and every rank renders its own camera (``rasterization(distributed=True)``), one checkpoint per rank as upstream.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from typing import Dict, List, Optional

import numpy as np
import torch
from torch import Tensor

SH_C0 = 0.2820947917738781


# ----------------------------------------------------------------------------------------------------------
# command line (the argument surface of gsplat's simple_trainer that the reference uses; dashes and underscores both work)
# ----------------------------------------------------------------------------------------------------------
def parse_args(argv: Optional[List[str]] = None) -> argparse.Namespace:
    argv = list(sys.argv[1:] if argv is None else argv)
    argv = [("--" + a[2:].replace("_", "-")) if a.startswith("--") else a for a in argv]
    ap = argparse.ArgumentParser(prog="simple_trainer.py", description=__doc__.split("\n\n")[0])
    ap.add_argument("strategy", choices=["default", "mcmc"])
    ap.add_argument("--data-dir", required=True)
    ap.add_argument("--result-dir", default="results")
    ap.add_argument("--data-factor", type=int, default=1)
    ap.add_argument("--max-steps", type=int, default=30_000)
    ap.add_argument("--steps-scaler", type=float, default=1.0)
    ap.add_argument("--batch-size", type=int, default=1)
    ap.add_argument("--packed", action="store_true")
    ap.add_argument("--disable-viewer", action="store_true")
    ap.add_argument("--sh-degree", type=int, default=3)
    ap.add_argument("--sh-degree-interval", type=int, default=1000)
    ap.add_argument("--ssim-lambda", type=float, default=0.2)
    ap.add_argument("--init-opa", type=float, default=None, help="default 0.1 (default strategy) / 0.5 (mcmc)")
    ap.add_argument("--init-scale", type=float, default=None, help="default 1.0 (default strategy) / 0.1 (mcmc)")
    ap.add_argument("--cap-max", type=int, default=1_000_000, help="mcmc: maximum number of Gaussians")
    ap.add_argument("--opacity-reg", type=float, default=None, help="default 0 / 0.01 (mcmc)")
    ap.add_argument("--scale-reg", type=float, default=None, help="default 0 / 0.01 (mcmc)")
    ap.add_argument("--antialiased", action="store_true")
    ap.add_argument("--visible-adam", action="store_true")
    ap.add_argument("--save-steps", type=int, nargs="*", default=[7_000, 30_000])
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--log-every", type=int, default=100)
    ap.add_argument("--dry-run", action="store_true",
                    help="parse the command line, read the COLMAP model, print the launch plan and exit 0 (no GPU needed)")
    a = ap.parse_args(argv)
    # gsplat's two presets (examples/simple_trainer.py: `default` and `mcmc` sub-commands)
    mcmc = a.strategy == "mcmc"
    a.init_opa = (0.5 if mcmc else 0.1) if a.init_opa is None else a.init_opa
    a.init_scale = (0.1 if mcmc else 1.0) if a.init_scale is None else a.init_scale
    a.opacity_reg = (0.01 if mcmc else 0.0) if a.opacity_reg is None else a.opacity_reg
    a.scale_reg = (0.01 if mcmc else 0.0) if a.scale_reg is None else a.scale_reg
    if a.batch_size < 1:
        raise ValueError("--batch-size must be >= 1")
    # gsplat scales every step-count option by steps_scaler (the reference passes 1 / n_gpus)
    s = a.steps_scaler
    a.max_steps = int(a.max_steps * s)
    a.save_steps = [int(x * s) for x in a.save_steps]
    a.sh_degree_interval = max(1, int(a.sh_degree_interval * s))
    return a


# ----------------------------------------------------------------------------------------------------------
# dataset
# ----------------------------------------------------------------------------------------------------------
class ColmapDataset:
    """images/ + sparse/0 of a COLMAP project; pinhole intrinsics (undistorted input, as the reference's SfM stage writes)."""

    def __init__(self, data_dir: str, factor: int = 1):
        import cv2
        from . import colmap
        sparse = colmap.find_sparse_dir(data_dir)
        cams = colmap.read_cameras(os.path.join(sparse, "cameras.bin"))
        imgs = colmap.read_images(os.path.join(sparse, "images.bin"))
        self.points, self.points_rgb = colmap.read_points3d(os.path.join(sparse, "points3D.bin"))
        img_dir = os.path.join(data_dir, "images" if factor == 1 else f"images_{factor}")
        if not os.path.isdir(img_dir):
            raise FileNotFoundError(img_dir)
        self._cv2 = cv2
        self.items = []
        for im in sorted(imgs.values(), key=lambda t: t.name):
            cam = cams[im.camera_id]
            if np.any(np.abs(cam.distortion()) > 1e-8):
                raise NotImplementedError(f"camera {cam.id} ({cam.model}) has lens distortion: undistort the images first")
            K = cam.K().copy()
            K[:2] /= factor
            self.items.append({"path": os.path.join(img_dir, im.name), "K": K.astype(np.float32),
                               "viewmat": im.world_to_cam().astype(np.float32),
                               "size": (cam.width // factor, cam.height // factor)})
        if not self.items:
            raise ValueError("the COLMAP model registers no image")
        centres = np.stack([np.linalg.inv(it["viewmat"])[:3, 3] for it in self.items])
        self.scene_scale = float(np.linalg.norm(centres - centres.mean(0), axis=1).max()) * 1.1 or 1.0
        self._cache: Dict[int, Tensor] = {}
        self._cache_bytes, self.cache_limit = 0, int(float(os.environ.get("B2S_IMAGE_CACHE_GB", "8")) * (1 << 30))

    def __len__(self):
        return len(self.items)

    def image_u8(self, i: int) -> Tensor:
        """RGB uint8 [H,W,3] on the host (pinned when a GPU is present).  The cache holds uint8 -- 12x smaller than the
        float image the loss consumes -- and is bounded (B2S_IMAGE_CACHE_GB per rank, default 8): beyond the bound
        images are decoded again on demand instead of page-locking tens of GB."""
        t = self._cache.get(i)
        if t is None:
            bgr = self._cv2.imread(self.items[i]["path"], self._cv2.IMREAD_COLOR)
            if bgr is None:
                raise FileNotFoundError(self.items[i]["path"])
            t = torch.from_numpy(np.ascontiguousarray(bgr[..., ::-1]))
            if self._cache_bytes + t.numel() <= self.cache_limit:
                if torch.cuda.is_available():
                    t = t.pin_memory()
                self._cache[i] = t
                self._cache_bytes += t.numel()
        return t

    def image(self, i: int, device=None) -> Tensor:
        """float32 [H,W,3] in [0,1]; converted on ``device`` when given (the uint8 bytes are what crosses PCIe)."""
        u8 = self.image_u8(i)
        if device is not None:
            return u8.to(device, non_blocking=True).float().div_(255.0)
        return u8.float().div_(255.0)


# ----------------------------------------------------------------------------------------------------------
# splats
# ----------------------------------------------------------------------------------------------------------
def knn_rms_dist(points: Tensor, k: int = 3, chunk: int = 4096) -> Tensor:
    """RMS distance to the k nearest neighbours, sqrt(mean(d^2)) -- upstream's initial scale (chunked brute force on the
    device; initialisation only)."""
    out = torch.empty(points.shape[0], device=points.device)
    for s in range(0, points.shape[0], chunk):
        d = torch.cdist(points[s:s + chunk], points)
        out[s:s + chunk] = d.topk(min(k + 1, points.shape[0]), largest=False).values[:, 1:].square().mean(dim=1).sqrt()
    return out


def create_splats(points: np.ndarray, rgb: np.ndarray, sh_degree: int, init_opa: float, init_scale: float, device,
                  rank: int = 0, world: int = 1) -> Dict[str, Tensor]:
    pts = torch.from_numpy(points).float().to(device)
    col = torch.from_numpy(rgb).float().to(device) / 255.0
    n_all = pts.shape[0]
    dist = knn_rms_dist(pts[: min(n_all, 200_000)]) if n_all > 1 else torch.ones(1, device=device)
    if n_all > dist.shape[0]:
        dist = torch.cat([dist, dist.median().expand(n_all - dist.shape[0])])
    pts, col, dist = pts[rank::world], col[rank::world], dist[rank::world]          # Gaussian sharding, as upstream
    N, K = pts.shape[0], (sh_degree + 1) ** 2
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    splats = {
        "means": pts.clone(),
        "scales": torch.log(dist.clamp_min(1e-7) * init_scale)[:, None].repeat(1, 3),
        "quats": torch.rand(N, 4, generator=g).to(device),
        "opacities": torch.logit(torch.full((N,), init_opa, device=device)),
        "sh0": ((col - 0.5) / SH_C0)[:, None, :].contiguous(),
        "shN": torch.zeros(N, K - 1, 3, device=device),
    }
    return {k: v.contiguous().requires_grad_() for k, v in splats.items()}


LRS = {"means": 1.6e-4, "scales": 5e-3, "quats": 1e-3, "opacities": 5e-2, "sh0": 2.5e-3, "shN": 2.5e-3 / 20}


# ----------------------------------------------------------------------------------------------------------
# densification: gsplat DefaultStrategy (clone / split / prune / opacity reset)
# ----------------------------------------------------------------------------------------------------------
class DefaultStrategy:
    def __init__(self, scene_scale: float, s: float = 1.0, prune_opa=0.005, grow_grad2d=0.0002, grow_scale3d=0.01,
                 prune_scale3d=0.1, refine_start=500, refine_stop=15_000, reset_every=3000, refine_every=100):
        self.scene_scale = scene_scale
        self.prune_opa, self.grow_grad2d, self.grow_scale3d, self.prune_scale3d = prune_opa, grow_grad2d, grow_scale3d, prune_scale3d
        self.refine_start, self.refine_stop = int(refine_start * s), int(refine_stop * s)
        self.reset_every, self.refine_every = max(1, int(reset_every * s)), max(1, int(refine_every * s))
        self.grad2d = self.count = None

    def _replace(self, splats, opt, fn_param, fn_state):
        for name in list(splats):
            old = splats[name]
            new = fn_param(name, old.detach()).contiguous().requires_grad_()
            for grp in opt.param_groups:
                for i, p in enumerate(grp["params"]):
                    if p is old:
                        st = opt.state.pop(old, None)
                        grp["params"][i] = new
                        if st is not None:
                            opt.state[new] = {"step": st["step"], "exp_avg": fn_state(st["exp_avg"]).contiguous(),
                                              "exp_avg_sq": fn_state(st["exp_avg_sq"]).contiguous()}
            splats[name] = new

    @torch.no_grad()
    def accumulate(self, splats, meta, step):
        if step >= self.refine_stop:
            return
        g = meta["means2d"].grad
        if g is None:
            return
        N = splats["means"].shape[0]
        if self.grad2d is None or self.grad2d.shape[0] != N:
            self.grad2d = torch.zeros(N, device=g.device)
            self.count = torch.zeros(N, device=g.device)
        g = g.clone()
        g[..., 0] *= meta["width"] / 2.0 * meta["n_cameras"]
        g[..., 1] *= meta["height"] / 2.0 * meta["n_cameras"]
        if meta.get("gaussian_ids") is not None:            # packed rows
            ids = meta["gaussian_ids"]
            off = getattr(self, "gid_offset", 0)
            ids = ids - off
            ok = (ids >= 0) & (ids < N)
            self.grad2d.index_add_(0, ids[ok], g.norm(dim=-1)[ok])
            self.count.index_add_(0, ids[ok], torch.ones_like(ids[ok], dtype=torch.float32))
        else:
            sel = meta["radii"] > 0                          # [C,N]
            self.grad2d += (g.norm(dim=-1) * sel).sum(0)
            self.count += sel.sum(0).float()

    @torch.no_grad()
    def refine(self, splats, opt, step) -> Dict[str, int]:
        stats = {}
        if step >= self.refine_stop or self.grad2d is None:
            return stats
        if step > self.refine_start and step % self.refine_every == 0:
            grads = self.grad2d / self.count.clamp_min(1)
            scales = torch.exp(splats["scales"].detach())
            high = grads > self.grow_grad2d
            small = scales.max(dim=-1).values <= self.grow_scale3d * self.scene_scale
            dup, split = high & small, high & ~small
            n_dup, n_split = int(dup.sum()), int(split.sum())
            if n_dup or n_split:
                quats = torch.nn.functional.normalize(splats["quats"].detach()[split], dim=-1)
                w, x, y, z = quats.unbind(-1)
                R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                                 2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                                 2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
                noise = torch.einsum("nij,nj,bnj->bni", R, scales[split], torch.randn(2, n_split, 3, device=scales.device))
                keep = ~split

                def fn_param(name, p):
                    parts = [p[keep], p[dup]]
                    if name == "means":
                        parts.append((p[split][None] + noise).reshape(-1, 3))
                    elif name == "scales":
                        parts.append(torch.log(torch.exp(p[split]) / 1.6).repeat(2, 1))
                    else:
                        parts.append(p[split].repeat(2, *([1] * (p.dim() - 1))))
                    return torch.cat(parts)

                def fn_state(v):
                    z = torch.zeros((n_dup + 2 * n_split,) + v.shape[1:], device=v.device)
                    return torch.cat([v[keep], z])

                self._replace(splats, opt, fn_param, fn_state)
            # prune
            opa = torch.sigmoid(splats["opacities"].detach())
            prune = opa < self.prune_opa
            if step > self.reset_every:
                prune |= torch.exp(splats["scales"].detach()).max(dim=-1).values > self.prune_scale3d * self.scene_scale
            n_prune = int(prune.sum())
            if n_prune:
                keep2 = ~prune
                self._replace(splats, opt, lambda name, p: p[keep2], lambda v: v[keep2])
            stats = {"duplicated": n_dup, "split": n_split, "pruned": n_prune, "n": splats["means"].shape[0]}
            self.grad2d = self.count = None
        if step % self.reset_every == 0 and step > 0:
            cap = math.log(2 * self.prune_opa / (1 - 2 * self.prune_opa))

            def fn_param(name, p):
                return p.clamp(max=cap) if name == "opacities" else p

            old_opa = splats["opacities"]
            self._replace(splats, opt, fn_param, lambda v: v)
            st = opt.state.get(splats["opacities"])
            if st is not None:
                st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
            del old_opa
            stats["opacity_reset"] = 1
        return stats


# ----------------------------------------------------------------------------------------------------------
# densification: gsplat MCMCStrategy ("3D Gaussian Splatting as Markov Chain Monte Carlo"): dead Gaussians are relocated
# onto live ones sampled by opacity, the population grows 5 % per refinement up to cap_max, and every step perturbs the
# means with covariance-shaped noise scaled by the means' learning rate
# ----------------------------------------------------------------------------------------------------------
def compute_relocation(opacities: Tensor, scales: Tensor, ratios: Tensor, n_max: int = 51):
    """Opacity and scale of the N copies that replace one Gaussian so that the rendered contribution is preserved
    (gsplat `compute_relocation`, eq. 9 of the paper): new_opacity = 1 - (1 - o)^(1/N);
    new_scale = scale * o / sum_{i=1..N} sum_{k=0..i-1} C(i-1, k) (-1)^k new_opacity^(k+1) / sqrt(k+1)."""
    N = ratios.clamp(1, n_max).to(torch.int64)
    o64 = opacities.double()
    new_o = 1.0 - torch.pow(1.0 - o64, 1.0 / N.double())
    denom = torch.zeros_like(o64)
    for i in range(1, int(N.max().item()) + 1 if N.numel() else 1):
        inner = torch.zeros_like(o64)
        for k in range(i):
            inner += math.comb(i - 1, k) * ((-1.0) ** k) * torch.pow(new_o, k + 1) / math.sqrt(k + 1)
        denom += torch.where(N >= i, inner, torch.zeros_like(inner))
    coeff = (o64 / denom.clamp_min(1e-300)).to(scales.dtype)
    return new_o.to(opacities.dtype), scales * coeff[:, None]


class MCMCStrategy(DefaultStrategy):
    def __init__(self, scene_scale: float, s: float = 1.0, cap_max=1_000_000, noise_lr=5e5, refine_start=500, refine_stop=25_000,
                 refine_every=100, min_opacity=0.005):
        self.scene_scale, self.cap_max, self.noise_lr, self.min_opacity = scene_scale, cap_max, noise_lr, min_opacity
        self.refine_start, self.refine_stop, self.refine_every = int(refine_start * s), int(refine_stop * s), max(1, int(refine_every * s))
        self.grad2d = self.count = None

    def accumulate(self, splats, meta, step):         # no gradient statistics in this strategy
        return

    def _resample(self, splats, idx, eps=torch.finfo(torch.float32).eps):
        """New opacity / scale of the sampled sources (each shared with its new copies)."""
        opa = torch.sigmoid(splats["opacities"].detach())
        ratios = torch.bincount(idx, minlength=opa.shape[0])[idx] + 1
        new_o, new_s = compute_relocation(opa[idx], torch.exp(splats["scales"].detach()[idx]), ratios)
        new_o = new_o.clamp(min=self.min_opacity, max=1.0 - eps)
        return torch.logit(new_o), torch.log(new_s)

    @torch.no_grad()
    def refine(self, splats, opt, step, lr: float = 0.0) -> Dict[str, int]:
        stats = {}
        if self.refine_start < step < self.refine_stop and step % self.refine_every == 0:
            opa = torch.sigmoid(splats["opacities"].detach())
            dead = opa <= self.min_opacity
            n_dead = int(dead.sum())
            if 0 < n_dead < opa.shape[0]:
                dead_idx, alive_idx = dead.nonzero(as_tuple=True)[0], (~dead).nonzero(as_tuple=True)[0]
                src = alive_idx[torch.multinomial(opa[alive_idx], n_dead, replacement=True)]
                new_logit, new_logs = self._resample(splats, src)

                def fn_param(name, p):
                    p = p.clone()
                    if name == "opacities":
                        p[src] = new_logit
                    elif name == "scales":
                        p[src] = new_logs
                    p[dead_idx] = p[src]
                    return p

                def fn_state(v):
                    v = v.clone()
                    v[src] = 0
                    v[dead_idx] = 0
                    return v

                self._replace(splats, opt, fn_param, fn_state)
            n = splats["means"].shape[0]
            n_new = max(0, min(self.cap_max, int(1.05 * n)) - n)
            if n_new > 0:
                opa = torch.sigmoid(splats["opacities"].detach())
                src = torch.multinomial(opa, n_new, replacement=True)
                new_logit, new_logs = self._resample(splats, src)

                def fn_param(name, p):
                    p = p.clone()
                    if name == "opacities":
                        p[src] = new_logit
                    elif name == "scales":
                        p[src] = new_logs
                    return torch.cat([p, p[src]])

                def fn_state(v):
                    v = v.clone()
                    v[src] = 0
                    return torch.cat([v, torch.zeros((n_new,) + v.shape[1:], device=v.device)])

                self._replace(splats, opt, fn_param, fn_state)
            stats = {"relocated": n_dead, "added": n_new, "n": splats["means"].shape[0]}
        # every step: covariance-shaped noise on the means, faded out on opaque Gaussians
        from .ops import quat_scale_to_covar_preci
        opa = torch.sigmoid(splats["opacities"].detach())
        covars, _ = quat_scale_to_covar_preci(splats["quats"].detach(), torch.exp(splats["scales"].detach()), compute_covar=True,
                                              compute_preci=False)
        gate = 1.0 / (1.0 + torch.exp(-100.0 * ((1.0 - opa) - 0.995)))
        noise = torch.randn_like(splats["means"]) * gate[:, None] * (lr * self.noise_lr)
        splats["means"].data.add_(torch.einsum("bij,bj->bi", covars, noise))
        return stats


# ----------------------------------------------------------------------------------------------------------
def optimizer_hparams(world: int, batch_size: int = 1):
    """Upstream's batch/world scaling (gsplat ``create_splats_with_optimizers``): with BS = batch_size * world views per
    optimizer step, lr * sqrt(BS), eps / sqrt(BS), betas = (1 - BS (1 - 0.9), 1 - BS (1 - 0.999)) -- this is what makes
    ``--steps_scaler 1/n`` on n GPUs (source/container/src/main.py:1321-1322) equivalent to the single-GPU schedule."""
    BS = max(1, int(batch_size) * int(world))
    return {"lr_scale": math.sqrt(BS), "eps": 1e-15 / math.sqrt(BS), "betas": (1.0 - BS * (1.0 - 0.9), 1.0 - BS * (1.0 - 0.999))}


def view_of_step(order: np.ndarray, step: int, rank: int, world: int, batch_size: int = 1, slot: int = 0) -> int:
    """One permutation shared by every rank (same seed), strided by rank as ``DistributedSampler`` does: in one step
    the ranks render ``world * batch_size`` different views (``slot`` = position inside this rank's batch)."""
    return int(order[((step * batch_size + slot) * world + rank) % len(order)])


def train(a: argparse.Namespace, rank: int = 0, world: int = 1, local_rank: int = 0) -> str:
    import torch.distributed as dist
    import pipeline_pointcloud_b200 as b2s
    from .training import FusedAdam, l1_ssim_loss

    if not torch.cuda.is_available():
        raise RuntimeError("simple_trainer needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(a.seed + rank)
    np.random.seed(a.seed + rank)
    order_rng = np.random.default_rng(a.seed)                 # the SAME view order on every rank

    data = ColmapDataset(a.data_dir, a.data_factor)
    splats = create_splats(data.points, data.points_rgb, a.sh_degree, a.init_opa, a.init_scale, dev, rank, world)
    hp = optimizer_hparams(world, a.batch_size)
    opt = FusedAdam([{"params": [splats[k]], "lr": LRS[k] * (data.scene_scale if k == "means" else 1.0) * hp["lr_scale"], "name": k}
                     for k in LRS], eps=hp["eps"], betas=hp["betas"])
    means_lr0 = opt.param_groups[0]["lr"]
    gamma = 0.01 ** (1.0 / max(a.max_steps, 1))
    strategy = (MCMCStrategy(data.scene_scale, a.steps_scaler, cap_max=a.cap_max) if a.strategy == "mcmc"
                else DefaultStrategy(data.scene_scale, a.steps_scaler))
    ckpt_dir = os.path.join(a.result_dir, "ckpts")
    os.makedirs(ckpt_dir, exist_ok=True)
    save_at = {s - 1 for s in a.save_steps} | {a.max_steps - 1}
    order = order_rng.permutation(len(data))
    epoch = 0
    last_ckpt, t0, stats = "", time.time(), {}
    strategy.gid_offset = 0
    for step in range(a.max_steps):
        B = a.batch_size
        e = (step * B * world) // len(order)                 # every rank crosses the epoch boundary at the same step
        if e != epoch:
            epoch, order = e, order_rng.permutation(len(data))
        views = [view_of_step(order, step, rank, world, B, b) for b in range(B)]      # --batch-size views per rank and step
        items = [data.items[i] for i in views]
        W, H = items[0]["size"]
        if any(tuple(it["size"]) != (W, H) for it in items):
            raise ValueError("--batch-size > 1 needs images of one size (as upstream's dataloader does)")
        target = torch.stack([data.image(i, dev) for i in views])
        deg = min(step // a.sh_degree_interval, a.sh_degree)
        colors = torch.cat([splats["sh0"], splats["shN"]], dim=1)
        render, alpha, meta = b2s.rasterization(
            splats["means"], splats["quats"], torch.exp(splats["scales"]), torch.sigmoid(splats["opacities"]), colors,
            torch.from_numpy(np.stack([it["viewmat"] for it in items])).to(dev), torch.from_numpy(np.stack([it["K"] for it in items])).to(dev),
            W, H, sh_degree=deg, packed=a.packed or world > 1, rasterize_mode="antialiased" if a.antialiased else "classic",
            distributed=world > 1)
        meta["means2d"].retain_grad()
        loss = l1_ssim_loss(render, target, a.ssim_lambda)
        if a.opacity_reg > 0.0:                         # mcmc preset: sparsity pressure on opacities and scales
            loss = loss + a.opacity_reg * torch.sigmoid(splats["opacities"]).abs().mean()
        if a.scale_reg > 0.0:
            loss = loss + a.scale_reg * torch.exp(splats["scales"]).abs().mean()
        opt.zero_grad()
        loss.backward()
        if world > 1 and (step == 0 or stats):       # shard sizes change only when densification ran (same step on every rank)
            n = torch.tensor([splats["means"].shape[0]], dtype=torch.int64, device=dev)
            ns = [torch.empty_like(n) for _ in range(world)]
            dist.all_gather(ns, n)
            strategy.gid_offset = int(sum(int(t) for t in ns[:rank]))
        # multi-GPU: this rank sees every rank's Gaussians through its own camera only; the statistics of its own
        # Gaussians therefore come from 1/world of the views (rows of other ranks' Gaussians are ignored)
        strategy.accumulate(splats, meta, step)
        vis = None
        if a.visible_adam and world == 1:
            vis = (meta["radii"] > 0).any(0) if meta.get("gaussian_ids") is None else \
                torch.zeros(splats["means"].shape[0], dtype=torch.bool, device=dev).index_fill_(0, meta["gaussian_ids"], True)
        opt.step(visibility=vis)
        opt.param_groups[0]["lr"] = means_lr0 * gamma ** (step + 1)
        stats = (strategy.refine(splats, opt, step, lr=opt.param_groups[0]["lr"]) if a.strategy == "mcmc"
                 else strategy.refine(splats, opt, step))
        if rank == 0 and (step % a.log_every == 0 or step == a.max_steps - 1 or stats):
            print(f"step {step:6d}  loss {loss.item():.5f}  N {splats['means'].shape[0]}  sh {deg}  "
                  f"{(time.time() - t0) / (step + 1) * 1e3:.1f} ms/it  {stats if stats else ''}", flush=True)
        if step in save_at:
            last_ckpt = os.path.join(ckpt_dir, f"ckpt_{step}_rank{rank}.pt")
            torch.save({"step": step, "splats": {k: v.detach().cpu() for k, v in splats.items()}}, last_ckpt)
    if world > 1:
        dist.barrier()
    return last_ckpt


# ----------------------------------------------------------------------------------------------------------
# launch: upstream's gsplat.distributed.cli, restated
# ----------------------------------------------------------------------------------------------------------
def launch_plan(env=None, n_local_gpus: Optional[int] = None) -> Dict[str, object]:
    """How this process must start the workers, from the environment the reference's runner leaves behind.

    * under torchrun (``TORCHELASTIC_RUN_ID`` / ``LOCAL_WORLD_SIZE`` set): the process IS a worker -- RANK / WORLD_SIZE /
      LOCAL_RANK are per-process and are used as they are;
    * otherwise (the reference's launch, source/container/src/main.py:620-655 and :1339-1347: plain ``python
      simple_trainer.py`` with ``CUDA_VISIBLE_DEVICES=0..n-1``): WORLD_SIZE counts *nodes* and RANK is the node rank, as in
      upstream's ``cli()``; one worker per visible GPU is spawned, world = nodes * local GPUs, rank = node * local + i.
    """
    env = os.environ if env is None else env
    if n_local_gpus is None:
        n_local_gpus = torch.cuda.device_count()
    if "TORCHELASTIC_RUN_ID" in env or "LOCAL_WORLD_SIZE" in env:
        return {"mode": "worker", "rank": int(env.get("RANK", "0")), "world": int(env.get("WORLD_SIZE", "1")),
                "local_rank": int(env.get("LOCAL_RANK", "0"))}
    nodes, node_rank = int(env.get("WORLD_SIZE", "1") or 1), int(env.get("RANK", "0") or 0)
    n_local = max(int(n_local_gpus), 1)
    world = nodes * n_local
    if world == 1:
        return {"mode": "single", "rank": 0, "world": 1, "local_rank": 0}
    addr = env.get("MASTER_ADDR", "127.0.0.1") or "127.0.0.1"
    if nodes == 1 and addr == "localhost":         # (main.py:648; a container's resolver may not know "localhost")
        addr = "127.0.0.1"
    return {"mode": "spawn", "nodes": nodes, "node_rank": node_rank, "n_local": n_local, "world": world,
            "master_addr": addr, "master_port": int(env.get("MASTER_PORT", "29500") or 29500)}


def _spawned_worker(local_rank: int, a: argparse.Namespace, plan: Dict[str, object]) -> None:
    import torch.distributed as dist
    rank = int(plan["node_rank"]) * int(plan["n_local"]) + local_rank
    world = int(plan["world"])
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": str(local_rank),
                       "LOCAL_WORLD_SIZE": str(plan["n_local"]), "MASTER_ADDR": str(plan["master_addr"]),
                       "MASTER_PORT": str(plan["master_port"])})
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", init_method=f"tcp://{plan['master_addr']}:{plan['master_port']}", rank=rank,
                            world_size=world, device_id=torch.device("cuda", local_rank))
    try:
        ckpt = train(a, rank, world, local_rank)
        if rank == 0:
            print(f"checkpoint: {ckpt}", flush=True)
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def launch(a: argparse.Namespace) -> int:
    plan = launch_plan()
    if a.dry_run or os.environ.get("B2S_TRAINER_DRY_RUN", "0") == "1":
        data = ColmapDataset(a.data_dir, a.data_factor)
        print(f"dry run: CUDA_VISIBLE_DEVICES={os.environ.get('CUDA_VISIBLE_DEVICES', '')} {plan}; {len(data)} views, {data.points.shape[0]} SfM points, max_steps {a.max_steps}, "
              f"save_steps {a.save_steps}, result_dir {a.result_dir}", flush=True)
        return 0
    if plan["mode"] == "spawn":
        if a.steps_scaler * plan["world"] > 1.0 + 1e-6 and plan["node_rank"] == 0:
            print(f"note: --steps_scaler {a.steps_scaler} with {plan['world']} ranks", flush=True)
        import torch.multiprocessing as mp
        # a worker that fails makes spawn raise -> non-zero exit code -> pipeline.py:226-235 stops the pipeline
        mp.spawn(_spawned_worker, args=(a, plan), nprocs=int(plan["n_local"]), join=True)
        return 0
    if plan["world"] == 1 and a.steps_scaler < 1.0:
        print(f"warning: --steps_scaler {a.steps_scaler} but only one GPU is visible: training {a.max_steps} steps on one rank",
              flush=True)
    ckpt = train(a, int(plan["rank"]), int(plan["world"]), int(plan["local_rank"]))
    if plan["rank"] == 0:
        print(f"checkpoint: {ckpt}")
    return 0


def main(argv: Optional[List[str]] = None) -> int:
    return launch(parse_args(argv))


if __name__ == "__main__":
    sys.exit(main())
