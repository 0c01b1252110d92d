"""Downstream fusion model and the slot-selection kernel on the GPU."""
import contextlib
import io
from functools import partial

import pytest
import torch
import torch.nn.functional as F

from oracle import devias_oracle as O
from util import assert_close, golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,S,C', [(8, 2, 101), (3, 4, 400), (1, 8, 101), (32, 2, 400)])
def test_slot_select_matches_torch(B, S, C):
    from devias_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(B * 5 + S)
    logits = torch.randn(B * S, C + 365, device='cuda', generator=g) * 3
    a, s = ops.slot_select(logits, S, C, 365)
    probs = F.softmax(logits, dim=-1).view(B, S, -1)
    ra = torch.argmax(probs[:, :, :C].max(dim=-1).values, dim=1)
    rs = torch.argmax(probs[:, :, C:C + 365].max(dim=-1).values, dim=1)
    assert a.dtype == torch.int64 and torch.equal(a, ra) and torch.equal(s, rs)


def test_slot_select_first_maximum_wins_on_ties():
    from devias_b200 import ops
    logits = torch.zeros(2 * 4, 101 + 365, device='cuda')        # every slot identical -> index 0
    a, s = ops.slot_select(logits, 4, 101, 365)
    assert a.tolist() == [0, 0] and s.tolist() == [0, 0]


def test_fusion_model_bf16_vs_golden_and_oracle():
    from devias_b200.modeling_slot_fusion import VisionTransformer
    g = golden('fusion_d2')
    sd = O.synth_fusion_state_dict(num_classes=101, depth=2, agg_depth=4, downstream_nb_classes=50, seed=8)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisionTransformer(patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=101, num_latents=2, head_type='mlp',
                              agg_weights_tie=True, agg_depth=4, slot_fusion_method='concat', downstream_nb_classes=50)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = O.synth_clips(2, seed=4)
    with torch.no_grad():
        inp, out = m(x.cuda())
        oinp, oout, _ = O.fusion_forward(sd, x, 101, depth=2)
    assert_close(inp, g['features'], 1e-2, 'fusion features vs reference golden')
    assert_close(out, g['logits'], 1e-2, 'fusion logits vs reference golden')
    assert_close(out, oout, 1e-2, 'fusion logits vs oracle')
    assert (out.argmax(-1).cpu().numpy() == g['logits'].argmax(-1)).all()


def test_fusion_model_trains():
    """gradients reach the fusion head, the norms, the aggregation block and the encoder"""
    from devias_b200.modeling_slot_fusion import VisionTransformer
    sd = O.synth_fusion_state_dict(num_classes=101, depth=2, agg_depth=4, downstream_nb_classes=50, seed=8)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisionTransformer(patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=101, num_latents=2, head_type='mlp',
                              agg_weights_tie=True, agg_depth=4, slot_fusion_method='concat', downstream_nb_classes=50)
    m.load_state_dict(sd)
    m = m.cuda().train()
    _, out = m(O.synth_clips(2, seed=4).cuda())
    F.cross_entropy(out, torch.tensor([3, 7], device='cuda')).backward()
    for k in ('fusion_head.classifier.weight', 'fusion_head.fc_action_down.weight', 'action_norm.weight', 'agg_block.latents',
              'blocks.0.attn.qkv.weight', 'patch_embed.proj.weight'):
        p = dict(m.named_parameters())[k]
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0, k
    assert dict(m.named_parameters())['fusion_head.fc_scene_down.weight'].grad is None      # unused in the reference too
