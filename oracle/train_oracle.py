"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the step either side of rasterization() in the reference's training
loop.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

The reference's Train-Stage1 component runs gsplat's `examples/simple_trainer.py`
(/root/reference/source/container/src/main.py:1328-1347); that script is not in /root/reference (gsplat is an unpinned
pip install, source/container/Dockerfile:220-229), so this restates its published behaviour with plain torch ops:

* loss = (1 - lambda) * mean|render - target| + lambda * (1 - SSIM(render, target)), lambda = 0.2, SSIM as fused_ssim /
  pytorch-msssim: 11x11 Gaussian window (sigma 1.5), "valid" padding, data range 1, C1 = 0.01^2, C2 = 0.03^2, mean over the map;
* torch.optim.Adam per parameter group (this IS the reference's optimizer: torch is what simple_trainer calls), with
  gsplat's SelectiveAdam variant: rows whose visibility mask is off keep parameters and moments untouched.

PARITY UNPINNED against gsplat itself (absent); pinned against torch.optim.Adam (the library the reference calls).
"""
import torch
import torch.nn.functional as F


def gaussian_window(size: int = 11, sigma: float = 1.5, dtype=torch.float32) -> torch.Tensor:
    x = torch.arange(size, dtype=torch.float64) - size // 2
    g = torch.exp(-(x * x) / (2.0 * sigma * sigma))
    return (g / g.sum()).to(dtype)


def ssim_map(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x, y: [C,H,W,ch] in [0,1]-ish; returns the SSIM map [C,ch,H-10,W-10] ("valid" padding)."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    x, y = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2)
    ch = x.shape[1]
    g = gaussian_window(dtype=x.dtype)
    w = (g[:, None] * g[None, :])[None, None].repeat(ch, 1, 1, 1)
    conv = lambda t: F.conv2d(t, w, groups=ch)
    mu1, mu2 = conv(x), conv(y)
    s1, s2, s12 = conv(x * x) - mu1 * mu1, conv(y * y) - mu2 * mu2, conv(x * y) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))


def l1_ssim_loss(render: torch.Tensor, target: torch.Tensor, ssim_lambda: float = 0.2) -> torch.Tensor:
    l1 = (render - target).abs().mean()
    return (1.0 - ssim_lambda) * l1 + ssim_lambda * (1.0 - ssim_map(render, target).mean())


def adam_reference(param, grads, lr, betas=(0.9, 0.999), eps=1e-8, masks=None):
    """Runs torch.optim.Adam over the list of gradients; with ``masks`` (per-row bool per step) applies gsplat's
    SelectiveAdam rule: a masked-off row keeps its parameter and both moments, the bias-correction step still counts."""
    p = param.clone()
    if masks is None:
        p.requires_grad_()
        opt = torch.optim.Adam([p], lr=lr, betas=betas, eps=eps)
        for g in grads:
            p.grad = g.clone()
            opt.step()
        return p.detach()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t, (g, mk) in enumerate(zip(grads, masks), start=1):
        mk = mk.reshape(-1, *([1] * (p.dim() - 1))).to(p.dtype)
        m_new = betas[0] * m + (1 - betas[0]) * g
        v_new = betas[1] * v + (1 - betas[1]) * g * g
        bc1, bc2 = 1 - betas[0] ** t, 1 - betas[1] ** t
        upd = (lr / bc1) * m_new / (v_new.sqrt() / (bc2 ** 0.5) + eps)
        p = p - mk * upd
        m, v = mk * m_new + (1 - mk) * m, mk * v_new + (1 - mk) * v
    return p
