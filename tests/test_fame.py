"""devias_b200.fame (device FAME) against the independent numpy restatement in oracle/fame_oracle.py.
Runs on the CPU (the module is plain torch); the GPU variant checks device placement and agreement with the CPU run."""
import numpy as np
import pytest
import torch

from oracle import fame_oracle as FO


def _clips(B, T=16, size=112, seed=0):
    rs = np.random.RandomState(seed)
    base = rs.uniform(0, 1, size=(B, 3, 1, size, size)).astype(np.float32)
    base = np.repeat(base, T, axis=2)
    base = np.clip(base + rs.normal(0, 0.02, size=base.shape).astype(np.float32), 0, 1)   # no exact ties between pixels
    for b in range(B):                      # a moving bright square = "foreground"
        for t in range(T):
            y = 20 + 3 * t + 5 * b
            base[b, :, t, y:y + 24, 30:60] = rs.uniform(0.6, 1.0, size=(3, 1, 1)).astype(np.float32)
    mean = np.array([0.485, 0.456, 0.406], np.float32).reshape(1, 3, 1, 1, 1)
    std = np.array([0.229, 0.224, 0.225], np.float32).reshape(1, 3, 1, 1, 1)
    return torch.from_numpy((base - mean) / std), base


def test_blur_and_hsv_match_independent_restatement():
    from devias_b200 import fame
    rs = np.random.RandomState(1)
    img = rs.uniform(size=(40, 52)).astype(np.float32)
    got = fame.gaussian_blur2d(torch.from_numpy(img)[None, None], 11, 11 / 3)[0, 0].numpy()
    assert np.abs(got - FO.gaussian_blur(img, 11, 11 / 3)).max() < 1e-6
    rgb = rs.uniform(size=(3, 33, 17)).astype(np.float32)
    rgb[:, 0, 0] = 0.5                       # grey pixel: delta = 0
    rgb[0, 1, 1] = rgb[1, 1, 1] = 0.9        # two channels share the maximum
    hsv = fame.rgb_to_hsv(torch.from_numpy(rgb)[None])[0].numpy()
    h, s, v = FO.rgb_to_hsv(rgb)
    assert np.abs(hsv[0] - h).max() < 1e-5 and np.abs(hsv[1] - s).max() < 1e-6 and np.abs(hsv[2] - v).max() < 1e-7


def test_masks_match_oracle_and_contract():
    from devias_b200.fame import FAME
    B = 3
    vids, denorm = _clips(B)
    f = FAME(crop_size=112, beta=0.5, prob_aug=1.0)
    tmp = vids * torch.tensor(f.frame_std).reshape(1, 3, 1, 1, 1) + torch.tensor(f.frame_mean).reshape(1, 3, 1, 1, 1)
    got = f.getmask(tmp).numpy()
    for b in range(B):
        diff = np.abs(denorm[b][:, :-1] - denorm[b][:, 1:]).sum(0).mean(0)
        ref = FO.binarise(FO.refine(FO.soft_mask(diff, f.gauss_size, f.gauss_sigma), denorm[b], f.gauss_size, f.gauss_sigma), 0.5)
        assert got[b].sum() == ref.sum() == int(0.5 * 112 * 112)
        assert (got[b] == ref).mean() > 0.995            # top-k ties on plateaus may fall differently
    torch.manual_seed(0)
    label = torch.arange(B)
    out_v, out_l, (m, mpf) = f(vids, label)
    assert out_v.shape == vids.shape and out_l.tolist() == [0, 1, 2]
    assert m.shape == (B, 49) and mpf.shape == (B, 8 * 49)          # 112 / 16 = 7 -> 49 cells per frame
    assert float(m.min()) >= 0 and float(m.max()) <= 1 and abs(float(m.mean()) - 0.5) < 1e-3
    torch.manual_seed(0)
    idx = torch.randperm(B)
    full = f.getmask(tmp)[:, None, None]
    assert torch.allclose(out_v, vids[idx] * (1 - full) + vids * full)


def test_prob_aug_branch_reorders_like_the_reference():
    from devias_b200.fame import FAME
    B = 4
    vids, _ = _clips(B, seed=3)
    f = FAME(crop_size=112, beta=0.5, prob_aug=0.5)
    torch.manual_seed(5)
    out_v, out_l, (m, mpf), cf = f(vids, torch.arange(B), center_frame=torch.arange(B) * 10)
    torch.manual_seed(5)
    torch.randperm(B)
    r = torch.rand(B)
    order = torch.cat([torch.where(r < 0.5)[0], torch.where(r >= 0.5)[0]])
    assert out_l.tolist() == order.tolist() and cf.tolist() == (order * 10).tolist()
    n_aug = int((r < 0.5).sum())
    assert torch.equal(out_v[n_aug:], vids[order[n_aug:]])          # the un-augmented samples pass through untouched
    assert m.shape == (B, 49) and mpf.shape == (B, 392)


@pytest.mark.gpu
def test_fame_on_device_matches_cpu():
    from devias_b200.fame import FAME
    vids, _ = _clips(2, size=224, seed=7)
    f = FAME(crop_size=224, beta=0.5, prob_aug=1.0)
    tmp = vids * torch.tensor(f.frame_std).reshape(1, 3, 1, 1, 1) + torch.tensor(f.frame_mean).reshape(1, 3, 1, 1, 1)
    cpu = f.getmask(tmp)
    gpu = f.getmask(tmp.cuda())
    assert gpu.is_cuda and (gpu.cpu() == cpu).float().mean() > 0.995
    out_v, out_l, (m, mpf) = f(vids.cuda(), torch.arange(2).cuda())
    assert out_v.is_cuda and m.is_cuda and m.shape == (2, 196) and mpf.shape == (2, 1568)


def test_blur_and_hsv_known_answers_from_the_published_definitions():
    """pins devias_b200.fame AND the numpy oracle on hand-computed known answers of the definitions the reference delegates to kornia
    for (tests/golden/fame_known_answers.json states the sources: kornia's docstring kernels, 'reflect' border, the HSV sextants)"""
    import json
    import os
    from devias_b200 import fame
    ka = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'fame_known_answers.json')))
    for case in ka['gaussian_kernel1d']:
        k = fame.gaussian_kernel1d(case['ksize'], case['sigma']).numpy()
        assert np.abs(k - np.array(case['kernel'])).max() < case['tol']
    for case in ka['blur_rows_k3_s2.5']:
        # a 5 x 5 image whose rows all equal `row`: the vertical pass of the separable blur leaves it unchanged (kernel sums to 1)
        img = np.tile(np.array(case['row'], np.float32), (5, 1))
        got = fame.gaussian_blur2d(torch.from_numpy(img)[None, None], 3, 2.5)[0, 0].numpy()
        ora = FO.gaussian_blur(img, 3, 2.5)
        for out in (got, ora):
            assert np.abs(out - np.array(case['out'])[None, :]).max() < 1e-4, (case, out[0])
    for case in ka['rgb_to_hsv']:
        rgb = np.array(case['rgb'], np.float32).reshape(3, 1, 1)
        got = fame.rgb_to_hsv(torch.from_numpy(rgb)[None])[0].numpy().reshape(3)
        h, s, v = FO.rgb_to_hsv(rgb)
        for out in (got, np.array([h.item(), s.item(), v.item()])):
            assert np.abs(out - np.array(case['hsv'])).max() < 2e-6, (case, out)
