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
# the configurations bench.py measures (BASELINE.json configs[1], configs[2]) at their full batch: (name, depth, S, agg_depth, tied, C, batch)
BENCH_CASES = [
    ('model_d12_ucf_b8', 12, 2, 4, True, 101, 8),
    ('model_d12_k400_b32', 12, 2, 8, True, 400, 32),
]
# train_class_batch with the REAL frozen teacher in the step (engine/engine_for_slot.py:50-56): (name, student depth, S, agg_depth, C, batch)
ENGINE_CASE = ('engine_teacher_d2', 2, 2, 4, 101, 4)


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


def val_targets(action_logit, seed):
    """Validation labels that make top-1 / top-5 non-trivial on random weights: a third of the clips are labelled with the
    reference's own arg-max (top-1 hit), a third with its 3rd-ranked class (top-5 hit only), the rest with its worst class."""
    order = torch.as_tensor(action_logit).argsort(dim=1, descending=True)
    rs = np.random.RandomState(seed)
    kind = rs.randint(0, 3, size=(order.shape[0],))
    pick = np.where(kind == 0, 0, np.where(kind == 1, 2, order.shape[1] - 1))
    return order[torch.arange(order.shape[0]), torch.from_numpy(pick)].to(torch.int64)


def val_metrics(output, target):
    """engine/engine_for_slot.py:234-239: CrossEntropyLoss over the unified C+365 row + timm.utils.accuracy(topk=(1, 5))
    (timm is absent: top-k restated -- prediction = 5 largest logits, hit = target among the first k, in percent)."""
    loss = torch.nn.functional.cross_entropy(output, target)
    pred = output.topk(5, 1, True, True).indices.t()
    correct = pred.eq(target.reshape(1, -1).expand_as(pred))
    acc1 = correct[:1].reshape(-1).float().sum(0) * 100.0 / output.shape[0]
    acc5 = correct[:5].reshape(-1).float().sum(0) * 100.0 / output.shape[0]
    return float(loss), float(acc1), float(acc5)


def engine_inputs(C, B, seed=21):
    rs = np.random.RandomState(seed)
    target = torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64))
    fg = torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32))
    fgf = torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32))
    return target, fg, fgf


def make_bench_cases(ns):
    """full-batch forwards of the benchmarked configurations (the reference evaluates clip by clip: chunks of 4 on the CPU)"""
    for name, depth, S, d, tied, C, B in BENCH_CASES:
        sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=3)
        m = ref_shim.build_student(ns, num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied)
        m.load_state_dict(sd); m.eval()
        x = O.synth_clips(B, seed=13)
        parts = []
        with torch.no_grad():
            for i in range(0, B, 4):
                parts.append(m(x[i:i + 4]))
        cat = lambda f: torch.cat([f(p) for p in parts], 0)
        al, sl = cat(lambda p: p[1][0]), cat(lambda p: p[1][1])
        attn = cat(lambda p: p[1][2])
        sh, slots, mp = cat(lambda p: p[2][0]), cat(lambda p: p[2][1]), cat(lambda p: p[2][2])
        af, sf = cat(lambda p: p[0][0]), cat(lambda p: p[0][1])
        tgt = val_targets(al, seed=17)
        loss, acc1, acc5 = val_metrics(al, tgt)
        np.savez_compressed(os.path.join(OUT, name + '.npz'), depth=depth, S=S, agg_depth=d, tied=tied, C=C, batch=B,
                            action_feat=_np(af), scene_feat=_np(sf), action_logit=_np(al), scene_logit=_np(sl),
                            attn_sample=_np(attn[..., ::SIM_STRIDE * 4]), attn_token_sum=_np(attn.sum(-1)),
                            slots_head=_np(sh), slots=_np(slots), mask_predictions=_np(mp),
                            val_target=tgt.numpy(), val_loss=np.float64(loss), val_acc1=np.float64(acc1), val_acc5=np.float64(acc5))
        print(name, float(al.abs().max()), loss, acc1, acc5)


def make_engine_case(ns):
    """engine/engine_for_slot.py:50-56 with the real teacher: student (train mode, drop rates 0) + frozen scene model under
    no_grad + the reference TrainLoss; stores the teacher logits, the total, its five parts and gradient norms of the total."""
    import contextlib, importlib, io
    name, depth, S, d, C, B = ENGINE_CASE
    sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=True, depth=depth, seed=14)
    m = ref_shim.build_student(ns, num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=True, depth=depth)
    m.load_state_dict(sd); m.train()
    mf = importlib.import_module('model.modeling_finetune')
    tsd = O.synth_teacher_state_dict(seed=15)
    with contextlib.redirect_stdout(io.StringIO()):
        tm = mf.vit_base_patch16_224(num_classes=365, use_mean_pooling=False, init_scale=1.0)
        crit = ns.train_loss.TrainLoss(torch.nn.CrossEntropyLoss(), 'KL', C, slot_matching_method='matching')
    tm.load_state_dict(tsd); tm.eval()
    x = O.synth_clips(B, seed=16)
    target, fg, fgf = engine_inputs(C, B)
    student_output = m(x)
    with torch.no_grad():
        teacher_output = tm(x, return_attn=False)
    total, act, parts = crit(m, student_output, teacher_output, target, fg_mask=(fg, fgf))
    # The reference objective cannot be differentiated in plain fp32 (its forced .half() masks make mse/bce backward raise
    # "Found dtype Half but expected Float"); gradients are therefore taken through the oracle's
    # restatement of the SAME objective evaluated on the reference model's own outputs, after checking the two values agree.
    ototal, _, _ = O.train_loss(student_output, teacher_output[1], target, (fg, fgf), C)
    assert abs(ototal.item() - total.item()) <= 1e-5 * abs(total.item()), (ototal.item(), total.item())
    ototal.backward()
    rec = {'total': np.float64(total.item()), 'teacher_logits': _np(teacher_output[1]), 'action_logit': _np(act)}
    for k, v in parts.items():
        rec['part/' + k] = np.float64(v)
    for k, p in m.named_parameters():
        rec['gnorm/' + k] = np.float64(p.grad.double().norm().item())
    np.savez_compressed(os.path.join(OUT, name + '.npz'), depth=depth, S=S, agg_depth=d, C=C, batch=B, **rec)
    print(name, total.item(), parts)


def main():
    ns = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if only in ('all', 'bench'):
        make_bench_cases(ns)
    if only in ('all', 'engine'):
        make_engine_case(ns)
    if only not in ('all', 'base'):
        return

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
