"""GPU: the fused training-objective kernels (csrc/train_loss.cu, one launch per direction) against the torch evaluation of the
same objective in devias_b200/loss.py -- which tests/test_loss.py pins on the reference's own TrainLoss values."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,S,C,crit', [(8, 2, 101, 'KL'), (5, 4, 400, 'KL'), (3, 2, 11, 'CE'), (32, 2, 400, 'KL'), (2, 8, 30, 'KL')])
def test_fused_train_loss_matches_torch_expressions(B, S, C, crit):
    from devias_b200 import _lib
    from devias_b200.loss import TrainLoss
    g = torch.Generator(device='cuda').manual_seed(B * 100 + S)
    W, H, N = C + 365, 4, 1568
    leaves = [torch.randn(B * S, W, device='cuda', generator=g) * 2.0,
              torch.rand(B * H, S, N, device='cuda', generator=g),
              torch.randn(B * S, 196, device='cuda', generator=g),
              torch.randn(B * S, 768, device='cuda', generator=g)]
    target = torch.randint(0, C, (B,), device='cuda', generator=g)
    teacher = torch.randn(B, 365, device='cuda', generator=g) * 3.0
    fg = (torch.rand(B, 196, device='cuda', generator=g) > 0.5).float()
    fgf = (torch.rand(B, N, device='cuda', generator=g) > 0.5).float()
    res = []
    for fused in (True, False):
        crit_ = TrainLoss(None, crit, C)
        crit_.fused = fused
        head, attn, maskp, slots = (t.clone().requires_grad_(True) for t in leaves)
        out = (None, (None, None, attn), (head, slots, maskp))
        n0 = _lib.launch_count()
        total, act, parts = crit_(None, out, (None, teacher), target, fg_mask=(fg, fgf))
        (total * 1.7).backward()
        launched = _lib.launch_count() - n0
        assert launched == (2 if fused else 0)
        res.append((total.detach(), act.detach(), {k: v.detach() if torch.is_tensor(v) else v for k, v in parts.items()},
                    [t.grad.clone() for t in (head, attn, maskp, slots)]))
    (t0, a0, p0, g0), (t1, a1, p1, g1) = res
    assert abs(float(t0) - float(t1)) <= 2e-5 * abs(float(t1)), (float(t0), float(t1))
    assert torch.equal(a0, a1), 'matched action rows differ (assignment)'
    for k in p1:
        assert abs(float(p0[k]) - float(p1[k])) <= 2e-5 * abs(float(p1[k])) + 1e-7, (k, float(p0[k]), float(p1[k]))
    for name, x, y in zip(('d slots_head', 'd attn', 'd mask_predictions', 'd slots'), g0, g1):
        err = float((x - y).norm() / (y.norm() + 1e-30))
        assert err <= 2e-5, (name, err)
