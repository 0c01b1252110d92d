"""GPU: the one-pass arena AdamW (csrc/optim.cu, devias_b200/optim.py) against torch.optim.AdamW -- per-group learning rates and
weight decays that change every step (the schedule the reference writes into param_groups, engine/engine_for_slot.py:91-97),
gradient clipping (clip_grad_norm_), the bf16 shadow and the in-pass gradient zero-fill."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(40, 72), torch.nn.GELU(), torch.nn.LayerNorm(72), torch.nn.Linear(72, 24),
                               torch.nn.Linear(24, 8, bias=False)).cuda()


def _groups(m):
    decay = [p for n, p in m.named_parameters() if p.dim() > 1]
    no_decay = [p for n, p in m.named_parameters() if p.dim() <= 1]
    return [dict(params=decay[:1], weight_decay=0.05, lr_scale=0.5), dict(params=decay[1:], weight_decay=0.05, lr_scale=1.0),
            dict(params=no_decay, weight_decay=0.0, lr_scale=1.0)]


@pytest.mark.parametrize('max_norm', [0.0, 0.05])
def test_arena_adamw_matches_torch_adamw(max_norm):
    from devias_b200.arena import ParamArena
    from devias_b200.optim import ArenaAdamW
    ref, ours = _net(), _net()
    ours.load_state_dict(copy.deepcopy(ref.state_dict()))
    arena = ParamArena.of(ours)
    o_ref = torch.optim.AdamW(_groups(ref), lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    o_our = ArenaAdamW(_groups(ours), arena, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, max_norm=max_norm)
    gen = torch.Generator(device='cuda').manual_seed(1)
    for it in range(6):
        lr = 1e-2 * (1.0 + 0.3 * it)                         # a schedule: every iteration rewrites lr / weight decay
        for opt in (o_ref, o_our):
            for g in opt.param_groups:
                g['lr'] = lr * g['lr_scale']
                if g['weight_decay'] > 0:
                    g['weight_decay'] = 0.05 + 0.01 * it
        x = torch.randn(16, 40, device='cuda', generator=gen)
        for m in (ref, ours):
            m(x).square().mean().backward()
        if max_norm > 0:
            total = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm)
        o_ref.step(); o_ref.zero_grad(set_to_none=True)
        o_our.step()
        if max_norm > 0:
            assert torch.allclose(o_our.grad_norm(), total, rtol=1e-4)
        assert float(arena.grad.abs().max()) == 0.0             # zero-filled in the same pass
        for (k, a), b in zip(ref.named_parameters(), ours.parameters()):
            # fp32 rounding of a differently ordered expression: a few 1e-5 of one update (lr ~ 1e-2)
            assert torch.allclose(a, b, rtol=2e-5, atol=3e-6), (it, k, float((a - b).abs().max()))
            assert b.grad is not None and b.grad.data_ptr() == arena.grad_view(b).data_ptr()
            assert torch.equal(arena.view16(b), b.detach().to(torch.bfloat16)), 'bf16 shadow not refreshed by the update pass'


def test_arena_rehomes_parameters_and_frozen_ones_stay_put():
    from devias_b200.arena import ParamArena
    from devias_b200.optim import ArenaAdamW
    m = _net()
    before = {k: v.detach().clone() for k, v in m.named_parameters()}
    arena = ParamArena.of(m)
    assert ParamArena.of(m) is arena
    for k, p in m.named_parameters():
        assert torch.equal(p, before[k]) and arena.contains(p)
    # parameters of the arena that are in no optimizer group are left untouched by the update pass
    params = list(m.parameters())
    opt = ArenaAdamW([dict(params=params[:2])], arena, lr=1e-2, weight_decay=0.1)
    m(torch.randn(4, 40, device='cuda')).sum().backward()
    opt.step()
    for p, (k, b) in zip(params, before.items()):
        changed = not torch.equal(p, b)
        assert changed == (p is params[0] or p is params[1]), k
