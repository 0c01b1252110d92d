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
