"""GPU parity of the fused loss and optimizer kernels against the torch restatement
(oracle/train_oracle.py; torch.optim.Adam is the very optimizer the reference's trainer calls)."""
import pytest
import torch

from oracle import train_oracle as TO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,H,W,ch", [(1, 64, 80, 3), (2, 37, 53, 3), (1, 27, 11, 1), (1, 200, 331, 4)])
def test_l1_ssim_loss_and_gradient(C, H, W, ch):
    from pipeline_pointcloud_b200.training import l1_ssim_loss
    g = torch.Generator().manual_seed(7)
    y = torch.rand(C, H, W, ch, generator=g)
    x = (y + 0.2 * torch.randn(C, H, W, ch, generator=g)).clamp(0, 1)
    x[0, : H // 3, : W // 4] = y[0, : H // 3, : W // 4]              # a region of exact equality: sign(0) = 0
    x64 = x.double().requires_grad_()
    ref = TO.l1_ssim_loss(x64, y.double())
    ref.backward()
    xg = x.cuda().requires_grad_()
    got = l1_ssim_loss(xg, y.cuda())
    got.backward()
    assert abs(got.item() - ref.item()) < 2e-6
    err = (xg.grad.cpu().double() - x64.grad).abs().max().item()
    assert err <= 1e-3 * x64.grad.abs().max().item() + 1e-9, err
    # forward only (no dmaps) gives the same value
    with torch.no_grad():
        assert abs(l1_ssim_loss(x.cuda(), y.cuda()).item() - got.item()) < 1e-7


def test_l1_ssim_rejects_small_images():
    from pipeline_pointcloud_b200._lib import B2SError
    from pipeline_pointcloud_b200.training import l1_ssim_loss
    with pytest.raises(B2SError):
        l1_ssim_loss(torch.rand(1, 10, 40, 3).cuda(), torch.rand(1, 10, 40, 3).cuda())


@pytest.mark.parametrize("shape", [(1000, 3), (257, 16, 3), (513,)])
def test_fused_adam_matches_torch_adam(shape):
    from pipeline_pointcloud_b200.training import FusedAdam
    g = torch.Generator().manual_seed(3)
    p0 = torch.randn(*shape, generator=g)
    grads = [torch.randn(*shape, generator=g) * (10.0 ** torch.randint(-4, 2, (1,), generator=g).item()) for _ in range(25)]
    ref = TO.adam_reference(p0, grads, lr=3e-3, betas=(0.9, 0.999), eps=1e-15)      # gsplat's eps
    p = p0.cuda().clone().requires_grad_()
    opt = FusedAdam([{"params": [p], "lr": 3e-3}], eps=1e-15)
    for gr in grads:
        p.grad = gr.cuda()
        opt.step()
    assert (p.detach().cpu() - ref).abs().max() <= 2e-6 * ref.abs().max() + 1e-7


def test_selective_adam_keeps_invisible_rows():
    from pipeline_pointcloud_b200.training import FusedAdam
    g = torch.Generator().manual_seed(4)
    p0 = torch.randn(400, 5, generator=g)
    grads = [torch.randn(400, 5, generator=g) for _ in range(12)]
    masks = [torch.rand(400, generator=g) > 0.4 for _ in range(12)]
    ref = TO.adam_reference(p0, grads, lr=1e-2, eps=1e-15, masks=masks)
    p = p0.cuda().clone().requires_grad_()
    opt = FusedAdam([p], lr=1e-2, eps=1e-15)
    for gr, mk in zip(grads, masks):
        p.grad = gr.cuda()
        opt.step(visibility=mk.cuda())
    assert (p.detach().cpu() - ref).abs().max() <= 2e-6 * ref.abs().max() + 1e-7
    never = ~torch.stack(masks).any(0)
    if never.any():
        assert torch.equal(p.detach().cpu()[never], p0[never])
