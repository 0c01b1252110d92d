"""CPU: the oracle restatement against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py) and, when /root/reference is present, against the reference live."""
import numpy as np
import pytest
import torch

from oracle import devias_oracle as O
from oracle import make_golden as MG
from oracle import ref_shim
from util import assert_close, golden

TOL = 2e-6  # fp32 restatement vs fp32 reference: same math, different op order only


@pytest.mark.parametrize('name,S,d,tied,B', MG.AGG_CASES)
def test_agg_block_golden(name, S, d, tied, B):
    g = golden(name)
    sd = {'agg_block.' + k: v for k, v in MG.agg_state(S, d, tied, seed=11).items()}
    x = O.synth_tokens(B, seed=5)
    with torch.no_grad():
        slots, sim = O.aggregation_block(sd, x)
    assert tuple(sim.shape) == tuple(g['sim_shape'])
    assert_close(slots, g['slots'], TOL, 'slots')
    assert_close(sim[..., ::MG.SIM_STRIDE], g['sim_sample'], TOL, 'sim sample')
    assert_close(sim.sum(-1), g['sim_token_sum'], TOL, 'sim token sums')
    # slot-axis softmax: columns sum to one (agg_block/attention.py:132)
    assert torch.allclose(sim.sum(1), torch.ones_like(sim.sum(1)), atol=1e-5)


@pytest.mark.parametrize('name,depth,S,d,tied,C,B', MG.MODEL_CASES)
def test_student_forward_golden(name, depth, S, d, tied, C, B):
    g = golden(name)
    sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=3)
    x = O.synth_clips(B, seed=1)
    with torch.no_grad():
        tokens = O.forward_features(sd, x)
        (af, sf), (al, sl, attn), (sh, slots, mp) = O.student_forward(sd, x, C)
    assert_close(tokens[:, ::97, ::5], g['tokens_sample'], TOL, 'tokens')
    assert_close(al, g['action_logit'], TOL, 'action_logit')
    assert_close(sl, g['scene_logit'], TOL, 'scene_logit')
    assert_close(af, g['action_feat'], TOL, 'action_feat')
    assert_close(sf, g['scene_feat'], TOL, 'scene_feat')
    assert_close(sh, g['slots_head'], TOL, 'slots_head')
    assert_close(slots, g['slots'], TOL, 'slots')
    assert_close(mp, g['mask_predictions'], TOL, 'mask_predictions')
    assert_close(attn[..., ::MG.SIM_STRIDE], g['attn_sample'], TOL, 'attn')
    assert (al[:, :C].argmax(-1).numpy() == g['action_logit'][:, :C].argmax(-1)).all()
    assert (sl[:, C:].argmax(-1).numpy() == g['scene_logit'][:, C:].argmax(-1)).all()


def test_grad_and_trainloss_golden():
    name, depth, S, d, tied, C, B = MG.GRAD_CASE
    g = golden(name)
    sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=4)
    # tied layers alias one tensor: differentiate w.r.t. unique storages
    uniq = {}
    for k, v in sd.items():
        uniq.setdefault(id(v), v.requires_grad_(True))
    out = O.student_forward(sd, O.synth_clips(B, seed=2), C)
    loss = MG.probe_loss(out)
    assert abs(loss.item() - float(g['probe_loss'])) <= 1e-4 * abs(float(g['probe_loss']))
    loss.backward()
    checked = 0
    for k, v in sd.items():
        key = 'gnorm/' + k
        if key not in g.files:
            continue  # aliases of tied layers (reference names them layers.0.*)
        gn = float(g[key])
        assert abs(v.grad.double().norm().item() - gn) <= 2e-4 * gn + 1e-6, k
        if gn > 1e-3:  # a few gradients are analytically zero (pure rounding noise): norm check only
            assert_close(v.grad.flatten()[:64], g['ghead/' + k], 5e-4, 'grad head ' + k)
        checked += 1
    assert checked >= 50
    rs = np.random.RandomState(9)
    target = torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64))
    teacher = torch.from_numpy(rs.standard_normal(size=(B, 365)).astype(np.float32))
    fg = torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32))
    fgf = torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32))
    with torch.no_grad():
        total, act, parts = O.train_loss(out, teacher, target, (fg, fgf), C)
    assert abs(float(total) - float(g['trainloss_total'])) <= 2e-5 * abs(float(g['trainloss_total']))
    for k, v in parts.items():
        assert abs(float(v) - float(g['trainloss/' + k])) <= 2e-5 * abs(float(g['trainloss/' + k])) + 1e-7, k
    assert_close(act, g['trainloss_action_logit'], TOL, 'trainloss action logits')


def test_sinusoid_known_answers():
    g = golden('sinusoid')
    tab = O.sinusoid_table(1568, 768)
    assert tab.shape == (1, 1568, 768)
    np.testing.assert_allclose(tab[0, [0, 1, 2, 777, 1567]].numpy(), g['rows'], rtol=0, atol=1e-7)
    # SURVEY.md section 8a row a4 probe values
    np.testing.assert_allclose(tab[0, 1, :4].numpy(), [0.8415, 0.5403, 0.8284, 0.5601], atol=5e-5)
    assert abs(tab.double().sum().item() - float(g['sum'])) < 1e-3


@pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present on this machine')
def test_oracle_vs_live_reference():
    ns = ref_shim.load()
    sd = O.synth_state_dict(num_classes=101, num_latents=4, agg_depth=3, agg_weights_tie=False, depth=1, seed=8)
    m = ref_shim.build_student(ns, num_latents=4, agg_depth=3, agg_weights_tie=False, depth=1)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd); m.eval()
    x = O.synth_clips(1, seed=6)
    with torch.no_grad():
        r = m(x)
        o = O.student_forward(sd, x, 101)
    assert_close(o[1][0], r[1][0], TOL, 'action_logit')
    assert_close(o[1][2], r[1][2], TOL, 'attn')
    assert_close(o[2][1], r[2][1], TOL, 'slots')
    assert_close(o[2][2], r[2][2], TOL, 'mask')


def test_teacher_forward_golden():
    g = golden('teacher_d12')
    sd = O.synth_teacher_state_dict(seed=6)
    with torch.no_grad():
        tok, logit = O.teacher_forward(sd, O.synth_clips(1, seed=3))
    assert_close(tok, g['token'], TOL, 'teacher token')
    assert_close(logit, g['logits'], TOL, 'teacher logits')
    assert int(logit.argmax()) == int(g['logits'].argmax())


def test_fusion_forward_golden():
    """downstream fusion model (model/modeling_slot_fusion.py): oracle restatement vs the reference's own output"""
    g = golden('fusion_d2')
    sd = O.synth_fusion_state_dict(num_classes=101, depth=2, agg_depth=4, downstream_nb_classes=50, seed=8)
    with torch.no_grad():
        inp, out, _ = O.fusion_forward(sd, O.synth_clips(2, seed=4), 101, depth=2)
    assert_close(inp, g['features'], TOL, 'fusion features')
    assert_close(out, g['logits'], TOL, 'fusion logits')
    assert (out.argmax(-1).numpy() == g['logits'].argmax(-1)).all()


def test_fusion_state_dict_keys_match_reference_layout():
    """the drop-in's module tree carries exactly the reference's parameter names and shapes (constructed on CPU)"""
    import contextlib, io
    from functools import partial
    from devias_b200.modeling_slot_fusion import VisionTransformer
    sd = O.synth_fusion_state_dict(num_classes=101, depth=2, agg_depth=4, downstream_nb_classes=50, seed=8)
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisionTransformer(patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=101, num_latents=2, head_type='mlp',
                              agg_weights_tie=True, agg_depth=4, slot_fusion_method='concat', downstream_nb_classes=50)
    own = m.state_dict()
    assert set(own.keys()) == set(sd.keys())
    for k, v in sd.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    m.load_state_dict(sd)
