"""GPU: the CUDA-graphed training step (single graph, and the three-graph data-parallel variant with its backward cut
point) produces the same parameter update as the eager step of engine.train_step."""
import contextlib
import copy
import io
from functools import partial

import numpy as np
import pytest
import torch

from oracle import devias_oracle as O

pytestmark = pytest.mark.gpu


def _model(C=11, depth=4):
    from devias_b200.modeling_slot import VisionTransformer
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisionTransformer(patch_size=16, embed_dim=768, depth=depth, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=C, num_latents=2, agg_depth=2,
                              agg_weights_tie=True, slot_matching_method='matching', init_scale=1.0)
    m.load_state_dict(O.synth_state_dict(num_classes=C, num_latents=2, agg_depth=2, depth=depth, seed=31))
    return m.cuda().train()


def _batch(C, B=2):
    rs = np.random.RandomState(4)
    return dict(clip=O.synth_clips(B, seed=8).cuda(),
                target=torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64)).cuda(),
                fg=torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32)).cuda(),
                fgf=torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32)).cuda(),
                teacher=torch.from_numpy(rs.standard_normal(size=(B, 365)).astype(np.float32)).cuda())


def _params(m):
    return {k: v.detach().clone() for k, v in m.named_parameters()}


def test_graphed_steps_match_eager():
    from devias_b200 import engine
    from devias_b200.ddp import GradReducer
    from devias_b200.loss import TrainLoss
    C = 11
    b = _batch(C)
    crit = TrainLoss(None, 'KL', C)
    results = []
    for mode in ('eager', 'graph', 'graph+reducer'):
        m = _model(C)
        opt = torch.optim.SGD(m.parameters(), lr=0.05)
        red = GradReducer(m) if mode == 'graph+reducer' else None
        if mode == 'eager':
            for _ in range(2):
                engine.train_step(m, None, crit, opt, b['clip'], b['target'], (b['fg'], b['fgf']), teacher_logits=b['teacher'])
        else:
            snap = copy.deepcopy(m.state_dict())
            step = engine.GraphedTrainStep(m, crit, opt, [b], reducer=red, warmup=1, split_block=2)
            m.load_state_dict(snap)          # undo the warm-up / capture updates, replay exactly two steps
            for _ in range(2):
                loss = step(0)
            assert torch.isfinite(loss)
        torch.cuda.synchronize()
        results.append(_params(m))
    ref = results[0]
    for other, name in zip(results[1:], ('graph', 'graph+reducer')):
        for k in ref:
            d = (other[k] - ref[k]).abs().max().item()
            scale = ref[k].abs().max().item() + 1e-6
            assert d <= 2e-3 * scale + 2e-5, (name, k, d, scale)


def _adamw_groups(m, lr):
    decay = [p for p in m.parameters() if p.dim() > 1]
    rest = [p for p in m.parameters() if p.dim() <= 1]
    return [dict(params=decay, weight_decay=0.05, lr=lr), dict(params=rest, weight_decay=0.0, lr=lr)]


@pytest.mark.parametrize('update_freq,max_norm', [(1, 0.0), (2, 0.5)])
def test_arena_graphed_step_follows_schedule_and_matches_eager_torch_adamw(update_freq, max_norm):
    """arena mode (ArenaAdamW + direct gradient accumulation, update in its own graph) against the eager step with
    torch.optim.AdamW: the learning rate changes EVERY optimizer step (the captured graph must follow it), gradients are
    accumulated over `update_freq` micro-batches and clipped to `max_norm` (engine/engine_for_slot.py:91-97, 147-166)."""
    from devias_b200 import engine
    from devias_b200.arena import ParamArena
    from devias_b200.loss import TrainLoss
    from devias_b200.optim import ArenaAdamW
    C = 11
    bs = [_batch(C), _batch(C)]
    bs[1]['clip'] = O.synth_clips(2, seed=9).cuda()
    crit = TrainLoss(None, 'KL', C)
    lrs = [2e-3, 1e-3, 3e-3]

    m = _model(C)
    opt = torch.optim.AdamW(_adamw_groups(m, lrs[0]), betas=(0.9, 0.999), eps=1e-8)
    micro = 0
    for lr in lrs:
        for g in opt.param_groups:
            g['lr'] = lr
        for u in range(update_freq):
            b = bs[micro % 2]; micro += 1
            engine.train_step(m, None, crit, opt, b['clip'], b['target'], (b['fg'], b['fgf']), teacher_logits=b['teacher'],
                              update_freq=update_freq, do_update=(u == update_freq - 1), max_norm=max_norm)
    ref = _params(m)

    m2 = _model(C)
    opt2 = ArenaAdamW(_adamw_groups(m2, lrs[0]), ParamArena.of(m2), betas=(0.9, 0.999), eps=1e-8, max_norm=max_norm)
    snap = copy.deepcopy(m2.state_dict())
    step = engine.GraphedTrainStep(m2, crit, opt2, bs, warmup=1, update_freq=update_freq)
    m2.load_state_dict(snap)
    opt2.exp_avg.zero_(); opt2.exp_avg_sq.zero_(); opt2._t = 0; opt2.arena.grad.zero_()
    micro = 0
    for lr in lrs:
        for g in opt2.param_groups:
            g['lr'] = lr
        for u in range(update_freq):
            loss = step(micro % 2); micro += 1
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    got = _params(m2)
    # compare the parameter UPDATES (Adam normalises them to ~lr per element, so following the lr schedule, the accumulation
    # and the clipping all show up in their size); the two paths run the same kernels, differences are atomics-order noise
    # amplified by Adam's normalisation of near-zero gradients
    du_ref = torch.cat([(ref[k] - snap[k].cuda()).flatten() for k in ref]).double()
    du_got = torch.cat([(got[k] - snap[k].cuda()).flatten() for k in ref]).double()
    rel = float((du_got - du_ref).norm() / du_ref.norm())
    worst = sorted(((float(((got[k] - ref[k]).double().norm()) / (float((ref[k] - snap[k].cuda()).double().norm()) + 1e-12)), k)
                    for k in ref), reverse=True)[:6]
    assert rel <= 0.05, (rel, worst)
    assert float(du_ref.abs().max()) > 0.5 * min(lrs)


def test_graphed_step_leaves_reducer_and_model_usable_for_eager_steps():
    """GraphedTrainStep must not leave the shared reducer switched off nor its cut-point hook active"""
    from devias_b200 import engine
    from devias_b200.ddp import GradReducer
    from devias_b200.loss import TrainLoss
    C = 11
    b = _batch(C)
    crit = TrainLoss(None, 'KL', C)
    m = _model(C)
    red = GradReducer(m)
    opt = torch.optim.SGD(m.parameters(), lr=0.01)
    step = engine.GraphedTrainStep(m, crit, opt, [b], reducer=red, warmup=1, cuts=[3, 1])
    assert len(step.graphs[0]) == 3                      # two cuts -> three backward pieces
    step(0)
    assert red.enabled is True
    step.close()
    assert not step._hooks
    before = _params(m)
    engine.train_step(m, None, crit, opt, b['clip'], b['target'], (b['fg'], b['fgf']), teacher_logits=b['teacher'], reducer=red)
    torch.cuda.synchronize()
    after = _params(m)
    assert any(not torch.equal(before[k], after[k]) for k in before)
    assert not step._taps
