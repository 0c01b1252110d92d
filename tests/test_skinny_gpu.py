"""Slot-side products (csrc/skinny.cu through devias_b200.slot_linear) against torch fp32/fp64 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize('M,K,N,bias', [(16, 768, 3072, True), (16, 3072, 768, True), (4, 768, 2048, False), (16, 768, 466, True),
                                        (32, 768, 765, True), (2, 256, 196, True), (37, 2048, 768, True)])
def test_linear_matches_torch(M, K, N, bias):
    from devias_b200 import slot_linear
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N)
    x = torch.randn(M, K, device='cuda', generator=g, requires_grad=True)
    w = (torch.randn(N, K, device='cuda', generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(N, device='cuda', generator=g).requires_grad_(True) if bias else None
    dy = torch.randn(M, N, device='cuda', generator=g)
    y = slot_linear.linear(x, w, b)
    grads = torch.autograd.grad(y, [x, w] + ([b] if bias else []), dy)
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    bd = b.detach().double().requires_grad_(True) if bias else None
    yr = F.linear(xd, wd, bd)
    gr = torch.autograd.grad(yr, [xd, wd] + ([bd] if bias else []), dy.double())
    assert _rel(y, yr) < 2e-6
    for a, r in zip(grads, gr):
        assert a.shape == r.shape and _rel(a, r) < 2e-6


@pytest.mark.parametrize('rows', [33, 64, 512, 1000])
def test_linear_3d_input_and_any_number_of_rows(rows):
    """no row limit and no library GEMM behind slot_linear: K400 B = 32 has 64 slot rows, an evaluation batch of 256 has 512"""
    from devias_b200 import _lib, slot_linear
    x = torch.randn(8, 2, 768, device='cuda')
    w = (torch.randn(512, 768, device='cuda') * 0.05).requires_grad_(True)
    b = torch.randn(512, device='cuda', requires_grad=True)
    assert _rel(slot_linear.linear(x, w), F.linear(x.double(), w.double())) < 2e-6
    xl = torch.randn(rows, 768, device='cuda', requires_grad=True)
    n0 = _lib.launch_count()
    y = slot_linear.linear(xl, w, b)
    dy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, [xl, w, b], dy)
    assert 3 <= _lib.launch_count() - n0 <= 4, 'forward (+ its pre-fill) and two backward kernels of csrc/skinny.cu'
    xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (xl, w, b))
    yr = F.linear(xd, wd, bd)
    rx, rw, rb = torch.autograd.grad(yr, [xd, wd, bd], dy.double())
    assert _rel(y, yr) < 2e-6 and _rel(gx, rx) < 2e-6 and _rel(gw, rw) < 2e-6 and _rel(gb, rb) < 2e-6


@pytest.mark.parametrize('B,S', [(8, 2), (1, 2), (3, 4), (32, 2), (5, 8)])
def test_fold_keys_and_apply_values(B, S):
    from devias_b200 import slot_linear
    H, dh, D = 4, 512, 768
    g = torch.Generator(device='cuda').manual_seed(B * 11 + S)
    q = torch.randn(B, S, H, dh, device='cuda', generator=g, requires_grad=True)
    wk = (torch.randn(H * dh, D, device='cuda', generator=g) * 0.05).requires_grad_(True)
    dqt = torch.randn(B, H, S, D, device='cuda', generator=g)
    qt = slot_linear.fold_keys(q, wk)
    gq, gw = torch.autograd.grad(qt, [q, wk], dqt)
    qd, wd = q.detach().double().requires_grad_(True), wk.detach().double().requires_grad_(True)
    ref = torch.einsum('bshd,hdc->bhsc', qd, wd.view(H, dh, D))
    rq, rw = torch.autograd.grad(ref, [qd, wd], dqt.double())
    assert _rel(qt, ref) < 2e-6 and _rel(gq, rq) < 2e-6 and _rel(gw, rw) < 2e-6

    cbar = torch.randn(B, H, S, D, device='cuda', generator=g, requires_grad=True)
    wv = (torch.randn(H * dh, D, device='cuda', generator=g) * 0.05).requires_grad_(True)
    dout = torch.randn(B, S, H * dh, device='cuda', generator=g)
    out = slot_linear.apply_values(cbar, wv)
    gc, gv = torch.autograd.grad(out, [cbar, wv], dout)
    cd, vd = cbar.detach().double().requires_grad_(True), wv.detach().double().requires_grad_(True)
    ref = torch.einsum('bhsc,hdc->bshd', cd, vd.view(H, dh, D)).reshape(B, S, H * dh)
    rc, rv = torch.autograd.grad(ref, [cd, vd], dout.double())
    assert _rel(out, ref) < 2e-6 and _rel(gc, rc) < 2e-6 and _rel(gv, rv) < 2e-6


@pytest.mark.parametrize('B,S', [(8, 2), (1, 2), (5, 4), (3, 8)])
def test_fold_epilogue_and_context_match_torch(B, S):
    """csrc/slot_glue.cu against the torch expressions of devias_b200/slot_attention.py, values and every gradient (fp64 reference)"""
    from devias_b200 import slot_linear
    H, D = 4, 768
    g_ = torch.Generator(device='cuda').manual_seed(B * 13 + S)
    rnd = lambda *s: torch.randn(*s, device='cuda', generator=g_)
    qt = rnd(B, H, S, D).requires_grad_(True)
    gamma = (1 + 0.1 * rnd(D)).requires_grad_(True)
    beta = (0.05 * rnd(D)).requires_grad_(True)
    scale = 512 ** -0.5
    g, G, c0 = slot_linear.fold_epilogue(qt, gamma, beta, scale)
    wg, wG, wc = rnd(B, H * S, D), rnd(B, H * S), rnd(B, H * S)
    grads = torch.autograd.grad((g * wg).sum() + (G * wG).sum() + (c0 * wc).sum(), [qt, gamma, beta])
    qd, gd, bd = (t.detach().double().requires_grad_(True) for t in (qt, gamma, beta))
    q2 = qd * scale
    rg = (q2 * gd).reshape(B, H * S, D); rG = rg.sum(-1); rc = (q2 @ bd).reshape(B, H * S)
    rgrads = torch.autograd.grad((rg * wg.double()).sum() + (rG * wG.double()).sum() + (rc * wc.double()).sum(), [qd, gd, bd])
    assert _rel(g, rg) < 2e-6 and _rel(G, rG) < 2e-6 and _rel(c0, rc) < 2e-6
    for a, r in zip(grads, rgrads):
        assert _rel(a, r) < 5e-6

    U = rnd(B, H * S, D).requires_grad_(True)
    m = rnd(B, H * S).requires_grad_(True)
    A = (rnd(B, H * S).abs() * 50 + 1).requires_grad_(True)
    cbar = slot_linear.context(U, m, A, gamma, beta, 1e-7)
    wcb = rnd(B, H * S, D)
    grads = torch.autograd.grad((cbar * wcb).sum(), [U, m, A, gamma, beta])
    Ud, md, Ad = (t.detach().double().requires_grad_(True) for t in (U, m, A))
    ref = (gd * (Ud - md.unsqueeze(-1)) + bd * Ad.unsqueeze(-1)) / (Ad.unsqueeze(-1) + 1e-7)
    rgrads = torch.autograd.grad((ref * wcb.double()).sum(), [Ud, md, Ad, gd, bd])
    assert _rel(cbar, ref) < 2e-6
    for a, r in zip(grads, rgrads):
        assert _rel(a, r) < 5e-6
