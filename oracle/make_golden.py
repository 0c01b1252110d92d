"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container (where /root/reference exists):  python -m oracle.make_golden
Inputs and weights are regenerated from seeds by oracle.devias_oracle.synth_* (numpy RandomState,
platform-stable), so only the reference's OUTPUTS are stored (full small tensors, strided samples of
the big ones plus their sums).  The reference has no golden vectors of its own;
these files are what pins the oracle -- and through it the CUDA path -- on machines without the tree.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import devias_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

# (name, S, agg_depth, tied, batch)
AGG_CASES = [
    ('agg_S2_d4_tied', 2, 4, True, 2),
    ('agg_S8_d3_tied', 8, 3, True, 2),
    ('agg_S4_d3_untied', 4, 3, False, 1),
    ('agg_S2_d8_tied', 2, 8, True, 1),
]
# (name, depth, S, agg_depth, tied, C, batch)
MODEL_CASES = [
    ('model_d12_ucf', 12, 2, 4, True, 101, 1),
    ('model_d2_k400', 2, 2, 8, True, 400, 2),
]
GRAD_CASE = ('grad_d2_ucf', 2, 2, 4, True, 101, 2)
SIM_STRIDE = 7


def _np(t):
    return t.detach().cpu().numpy().astype(np.float32)


def agg_state(S, depth, tied, seed):
    sd = O.synth_state_dict(num_latents=S, agg_depth=depth, agg_weights_tie=tied, depth=0, seed=seed)
    return {k[len('agg_block.'):]: v for k, v in sd.items() if k.startswith('agg_block.')}


def probe_weights(shapes, seed):
    rs = np.random.RandomState(seed)
    return [torch.from_numpy(rs.standard_normal(size=s).astype(np.float32)) for s in shapes]


def probe_loss(out, seed=77):
    """A fixed random linear functional of every differentiable student output (so that every
    parameter, incl. the last layer's slot softmax `attn`, receives gradient)."""
    (af, sf), (al, sl, attn), (sh, slots, mp) = out
    ts = [af, sf, al, sl, attn, sh, slots, mp]
    ws = probe_weights([tuple(t.shape) for t in ts], seed)
    return sum((t * w).sum() for t, w in zip(ts, ws))


def main():
    ns = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    for name, S, d, tied, B in AGG_CASES:
        sd = agg_state(S, d, tied, seed=11)
        with torch.no_grad():
            import contextlib, io
            with contextlib.redirect_stdout(io.StringIO()):
                m = ns.agg_block.AggregationBlock(num_latents=S, weight_tie_layers=tied, depth=d)
            m.load_state_dict(sd); m.eval()
            x = O.synth_tokens(B, seed=5)
            slots, sim = m(x)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), S=S, depth=d, tied=tied, batch=B,
                            slots=_np(slots), sim_sample=_np(sim[..., ::SIM_STRIDE]),
                            sim_token_sum=_np(sim.sum(-1)), sim_shape=np.array(sim.shape))
        print(name, tuple(slots.shape), tuple(sim.shape))

    for name, depth, S, d, tied, C, B in MODEL_CASES:
        sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=3)
        m = ref_shim.build_student(ns, num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth)
        m.load_state_dict(sd); m.eval()
        x = O.synth_clips(B, seed=1)
        with torch.no_grad():
            tokens = m.forward_features(x)
            (af, sf), (al, sl, attn), (sh, slots, mp) = m(x)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), depth=depth, S=S, agg_depth=d, tied=tied, C=C, batch=B,
                            tokens_sample=_np(tokens[:, ::97, ::5]), tokens_sum=_np(tokens.sum((1, 2))),
                            action_feat=_np(af), scene_feat=_np(sf), action_logit=_np(al), scene_logit=_np(sl),
                            attn_sample=_np(attn[..., ::SIM_STRIDE]), attn_token_sum=_np(attn.sum(-1)),
                            slots_head=_np(sh), slots=_np(slots), mask_predictions=_np(mp))
        print(name, float(al.abs().max()))

    # gradients of a fixed linear probe + the reference TrainLoss forward values
    name, depth, S, d, tied, C, B = GRAD_CASE
    sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=4)
    m = ref_shim.build_student(ns, num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth)
    m.load_state_dict(sd); m.train()  # drop rates are 0 -> deterministic
    x = O.synth_clips(B, seed=2)
    out = m(x)
    loss = probe_loss(out)
    loss.backward()
    rec = {'probe_loss': np.float64(loss.item())}
    seen = {}
    for k, p in m.named_parameters():  # tied layers appear once (layers.0.*)
        g = p.grad
        rec['gsum/' + k] = np.float64(g.double().sum().item())
        rec['gnorm/' + k] = np.float64(g.double().norm().item())
        rec['ghead/' + k] = _np(g.flatten()[:64])
        seen[k] = True
    # reference TrainLoss forward (binary masks => its .half() casts are exact)
    rs = np.random.RandomState(9)
    target = torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64))
    teacher = torch.from_numpy(rs.standard_normal(size=(B, 365)).astype(np.float32))
    fg = torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32))
    fgf = torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32))
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        crit = ns.train_loss.TrainLoss(torch.nn.CrossEntropyLoss(), 'KL', C, slot_matching_method='matching')
    with torch.no_grad():
        total, act, parts = crit(m, out, (None, teacher), target, fg_mask=(fg, fgf))
    rec['trainloss_total'] = np.float64(total.item())
    for k, v in parts.items():
        rec['trainloss/' + k] = np.float64(v)
    rec['trainloss_action_logit'] = _np(act)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), depth=depth, S=S, agg_depth=d, tied=tied, C=C, batch=B, **rec)
    print(name, loss.item(), total.item(), parts)

    # frozen scene teacher (model/modeling_finetune.py), CLS token, 365 classes
    import importlib
    mf = importlib.import_module('model.modeling_finetune')
    tsd = O.synth_teacher_state_dict(seed=6)
    with contextlib.redirect_stdout(io.StringIO()):
        tm = mf.vit_base_patch16_224(num_classes=365, use_mean_pooling=False, init_scale=1.0)
    assert set(tm.state_dict().keys()) == set(tsd.keys())
    tm.load_state_dict(tsd); tm.eval()
    with torch.no_grad():
        tok, logit = tm(O.synth_clips(1, seed=3))
    np.savez_compressed(os.path.join(OUT, 'teacher_d12.npz'), token=_np(tok), logits=_np(logit))
    print('teacher', float(logit.abs().max()))

    # downstream fusion model (model/modeling_slot_fusion.py), head_type='mlp', concat
    msf = importlib.import_module('model.modeling_slot_fusion')
    fsd = O.synth_fusion_state_dict(num_classes=101, depth=2, agg_depth=4, downstream_nb_classes=50, seed=8)
    from functools import partial
    with contextlib.redirect_stdout(io.StringIO()):
        fm = msf.VisionTransformer(patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                                   norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=101, num_latents=2,
                                   head_type='mlp', agg_weights_tie=True, agg_depth=4, slot_fusion_method='concat',
                                   downstream_nb_classes=50, use_input_ln=True)
    assert set(fm.state_dict().keys()) == set(fsd.keys()), set(fm.state_dict().keys()) ^ set(fsd.keys())
    fm.load_state_dict(fsd); fm.eval()
    with torch.no_grad():
        finp, fout = fm(O.synth_clips(2, seed=4))
    np.savez_compressed(os.path.join(OUT, 'fusion_d2.npz'), features=_np(finp), logits=_np(fout))
    print('fusion', float(fout.abs().max()))

    # sinusoid table known answers
    tab = ns.modeling_slot.get_sinusoid_encoding_table(1568, 768)
    np.savez_compressed(os.path.join(OUT, 'sinusoid.npz'), rows=_np(tab[0, [0, 1, 2, 777, 1567]]), sum=np.float64(tab.double().sum().item()))


if __name__ == '__main__':
    main()
