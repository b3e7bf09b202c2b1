"""The two fused kernels either side of ``rasterization()`` in a training step: the photometric
loss that produces the render cotangent and the Adam update that consumes the gradients.

Reference anchor: the Train-Stage1 component runs gsplat's ``examples/simple_trainer.py``
(source/container/src/main.py:1328-1347): ``loss = 0.8 * L1 + 0.2 * (1 - fused_ssim(render, target, padding="valid"))``
and one ``torch.optim.Adam`` (``SelectiveAdam`` under ``--packed``/``--visible_adam``) per parameter group.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
from torch import Tensor

from . import _lib as L
from ._lib import check, f32c, ptr


class _L1SSIM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render, target, ssim_lambda):
        dev = L.require_cuda(render, target)
        render, target = f32c(render), f32c(target)
        if render.shape != target.shape or render.dim() != 4:
            raise L.B2SError("render and target must both be [C,H,W,channels]")
        C, H, W, ch = render.shape
        lib, st = L.lib(), L.stream_ptr(dev)
        n_ctas = int(lib.b2s_l1_ssim_ctas(C, H, W))
        if n_ctas == 0:
            raise L.B2SError("images must be larger than the 11x11 SSIM window")
        need_grad = render.requires_grad
        partial = torch.empty(n_ctas, 2, device=dev)
        dmaps = torch.empty(C, H - 10, W - 10, ch, 3, device=dev) if need_grad else None
        check(lib.b2s_l1_ssim_fwd(C, H, W, ch, ptr(render), ptr(target), ptr(dmaps), ptr(partial), st), "b2s_l1_ssim_fwd")
        sums = partial.sum(dim=0)
        n_l1, n_ss = C * H * W * ch, C * (H - 10) * (W - 10) * ch
        loss = (1.0 - ssim_lambda) * sums[0] / n_l1 + ssim_lambda * (1.0 - sums[1] / n_ss)
        ctx.save_for_backward(render, target, dmaps)
        ctx.cfg = (float(ssim_lambda), n_l1, n_ss)
        return loss

    @staticmethod
    def backward(ctx, v):
        render, target, dmaps = ctx.saved_tensors
        lam, n_l1, n_ss = ctx.cfg
        C, H, W, ch = render.shape
        v = float(v)          # the loss is the root of the graph in a trainer (v == 1); a scalar multiple is folded in
        out = torch.empty_like(render)
        check(L.lib().b2s_l1_ssim_bwd(C, H, W, ch, ptr(render), ptr(target), ptr(dmaps), v * (1.0 - lam) / n_l1, -v * lam / n_ss,
                                      ptr(out), L.stream_ptr(render.device)), "b2s_l1_ssim_bwd")
        return out, None, None


def l1_ssim_loss(render: Tensor, target: Tensor, ssim_lambda: float = 0.2) -> Tensor:
    """``(1 - lambda) * L1 + lambda * (1 - SSIM)`` of [C,H,W,ch] images in one forward and one backward kernel."""
    return _L1SSIM.apply(render, target, float(ssim_lambda))


class FusedAdam:
    """``torch.optim.Adam`` semantics (no weight decay, no amsgrad) with one fused kernel per parameter tensor;
    ``step(visibility=mask)`` is gsplat's SelectiveAdam: rows whose mask is off keep parameters and moments.
    ``param_groups`` / ``state`` / ``state_dict`` mirror torch's so densification strategies can edit the moments."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = list(g["params"]) if not isinstance(g["params"], Tensor) else [g["params"]]
            g.setdefault("lr", lr); g.setdefault("betas", betas); g.setdefault("eps", eps)
            self.param_groups.append(g)
        self.state: Dict[Tensor, dict] = {}

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
            self.state[p] = st
        return st

    @torch.no_grad()
    def step(self, visibility: Optional[Tensor] = None):
        lib = L.lib()
        mask = visibility.to(torch.uint8).contiguous() if visibility is not None else None
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                L.require_cuda(p)
                if not p.is_contiguous() or p.dtype != torch.float32:
                    raise L.B2SError("FusedAdam needs contiguous fp32 parameters")
                st = self._state(p)
                st["step"] += 1
                rows = p.shape[0] if p.dim() > 0 else 1
                width = p.numel() // max(rows, 1)
                if mask is not None and mask.numel() != rows:
                    raise L.B2SError("visibility mask must have one entry per row")
                check(lib.b2s_adam_step(rows, width, ptr(p), ptr(f32c(p.grad)), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]),
                                        ptr(mask), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                        int(st["step"]), L.stream_ptr(p.device)), "b2s_adam_step")

    def state_dict(self):
        return {"param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                "state": [[{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._state(p).items()} for p in g["params"]]
                          for g in self.param_groups]}
