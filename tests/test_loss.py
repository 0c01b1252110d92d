"""CPU: the batched device-side TrainLoss (devias_b200/loss.py) against the reference's TrainLoss values (golden) and
the oracle restatement -- host logic, no GPU needed."""
import numpy as np
import torch

from devias_b200.loss import TrainLoss
from oracle import devias_oracle as O
from oracle import make_golden as MG
from util import assert_close, golden


def _inputs(C, B):
    rs = np.random.RandomState(9)
    target = torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64))
    teacher = torch.from_numpy(rs.standard_normal(size=(B, 365)).astype(np.float32))
    fg = torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32))
    fgf = torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32))
    return target, teacher, fg, fgf


def test_trainloss_matches_reference_golden():
    name, depth, S, d, tied, C, B = MG.GRAD_CASE
    g = golden(name)
    sd = O.synth_state_dict(num_classes=C, num_latents=S, agg_depth=d, agg_weights_tie=tied, depth=depth, seed=4)
    with torch.no_grad():
        out = O.student_forward(sd, O.synth_clips(B, seed=2), C)
    target, teacher, fg, fgf = _inputs(C, B)
    crit = TrainLoss(torch.nn.CrossEntropyLoss(), 'KL', C)
    total, act, parts = crit(None, out, (None, teacher), target, fg_mask=(fg, fgf))
    assert abs(float(total) - float(g['trainloss_total'])) <= 2e-5 * abs(float(g['trainloss_total']))
    for k, v in parts.items():
        assert abs(float(v) - float(g['trainloss/' + k])) <= 2e-5 * abs(float(g['trainloss/' + k])) + 1e-7, k
    assert_close(act, g['trainloss_action_logit'], 2e-6, 'action rows')


def test_matching_equals_bruteforce_many_slots():
    rs = np.random.RandomState(3)
    for S in (2, 3, 5, 8):
        B, C = 16, 30
        head = torch.from_numpy(rs.standard_normal(size=(B * S, C + 365)).astype(np.float32))
        target = torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64))
        scene_target = torch.from_numpy(rs.randint(C, C + 365, size=(B,)).astype(np.int64))
        sfm = head.softmax(-1)
        ai, si = TrainLoss.match(sfm.view(B, S, -1), target, scene_target)
        for b in range(B):
            cost = torch.stack([-sfm[b * S:(b + 1) * S, target[b]], -sfm[b * S:(b + 1) * S, scene_target[b]]], dim=1)
            assert O._assign(cost) == (int(ai[b]), int(si[b]))
        try:
            from scipy.optimize import linear_sum_assignment
        except Exception:
            continue
        for b in range(B):
            cost = torch.stack([-sfm[b * S:(b + 1) * S, target[b]], -sfm[b * S:(b + 1) * S, scene_target[b]]], dim=1)
            r, c = linear_sum_assignment(cost.numpy())
            got = {int(cc): int(rr) for rr, cc in zip(r, c)}
            assert got[0] == int(ai[b]) and got[1] == int(si[b])


def test_loss_gradient_matches_oracle():
    C, B, S = 20, 3, 2
    rs = np.random.RandomState(5)
    mk = lambda *s: torch.from_numpy(rs.standard_normal(size=s).astype(np.float32))
    head, slots, mp, attn = mk(B * S, C + 365), mk(B * S, 768), mk(B * S, 196), torch.rand(B * 4, S, 1568)
    target, teacher, fg, fgf = _inputs(C, B)
    outs = []
    for fn in ('mine', 'oracle'):
        ts = [t.clone().requires_grad_(True) for t in (head, slots, mp, attn)]
        so = ((None, None), (None, None, ts[3]), (ts[0], ts[1], ts[2]))
        if fn == 'mine':
            total, _, _ = TrainLoss(None, 'KL', C)(None, so, (None, teacher), target, fg_mask=(fg, fgf))
        else:
            total, _, _ = O.train_loss(so, teacher, target, (fg, fgf), C)
        total.backward()
        outs.append((total.detach(), [t.grad for t in ts]))
    assert abs(float(outs[0][0]) - float(outs[1][0])) <= 1e-5 * abs(float(outs[1][0]))
    for a, b in zip(outs[0][1], outs[1][1]):
        assert_close(a, b, 1e-5, 'loss grads')
