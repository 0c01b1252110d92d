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
def test_arena_graphed_step_gradients_accumulation_clipping_and_schedule(update_freq, max_norm):
    """arena mode (ArenaAdamW + direct gradient accumulation, update in its own graph).  Adam's normalised update amplifies
    bf16 run-to-run noise (two eager torch runs of this step differ by 60 % in their updates), so the checks are made on
    quantities that are LINEAR in the gradient:
      * with lr = 0 one captured optimizer step leaves exp_avg = (1 - beta1) * clip * sum of the micro-batch gradients
        (engine/engine_for_slot.py:147-166: loss / update_freq, clip_grad_norm_) -- compared with eager autograd gradients;
      * the captured update reads lr from device memory on every replay (engine_for_slot.py:91-97): the first Adam step has
        size lr per element, so halving lr between two replays from the same state halves the update."""
    from devias_b200 import engine
    from devias_b200.arena import ParamArena
    from devias_b200.loss import TrainLoss
    from devias_b200.optim import ArenaAdamW
    C = 11
    bs = [_batch(C), _batch(C)]
    bs[1]['clip'] = O.synth_clips(2, seed=9).cuda()
    crit = TrainLoss(None, 'KL', C)

    # eager autograd reference of the accumulated gradient
    m = _model(C)
    for u in range(update_freq):
        b = bs[u % 2]
        loss, _, _ = engine.train_class_batch(m, None, b['clip'], b['target'], crit, (b['fg'], b['fgf']), teacher_logits=b['teacher'])
        (loss / update_freq).backward()
    gref = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    total = torch.sqrt(sum(g.double().square().sum() for g in gref.values()))
    coef = min(1.0, max_norm / (float(total) + 1e-6)) if max_norm > 0 else 1.0

    m2 = _model(C)
    opt = ArenaAdamW(_adamw_groups(m2, 1e-3), ParamArena.of(m2), betas=(0.9, 0.999), eps=1e-8, max_norm=max_norm)
    snap = {k: v.detach().clone() for k, v in m2.state_dict().items()}
    step = engine.GraphedTrainStep(m2, crit, opt, bs, warmup=1, update_freq=update_freq)

    def reset(lr, wd):
        m2.load_state_dict(snap)
        opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt._t = 0; opt.arena.grad.zero_()
        step._micro = 0
        for g in opt.param_groups:
            g['lr'], g['weight_decay'] = lr, wd

    def one_update():
        for u in range(update_freq):
            loss = step(u % 2)
        torch.cuda.synchronize()
        assert torch.isfinite(loss)

    reset(0.0, 0.0)
    one_update()
    for k, p in m2.named_parameters():
        assert torch.equal(p.detach(), snap[k]), f'{k} moved although lr = 0'
        want = 0.1 * coef * gref[k].double()
        got = opt.state[p]['exp_avg'].double()
        if float(want.norm()) < 1e-7 * float(total):
            continue                                   # analytically zero gradients (noise only)
        err = float((got - want).norm() / want.norm())
        assert err <= 3e-2, (k, err)
    assert float(opt.arena.grad.abs().max()) == 0.0    # the update pass zero-filled the arena for the next step
    if max_norm > 0:
        assert abs(float(opt.grad_norm()) - float(total)) <= 2e-2 * float(total)

    sizes = []
    for lr in (2e-3, 1e-3):
        reset(lr, 0.0)
        one_update()
        sizes.append(float(torch.cat([(p.detach() - snap[k]).flatten() for k, p in m2.named_parameters()]).double().norm()))
    assert abs(sizes[0] / sizes[1] - 2.0) <= 0.04, sizes


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
