"""Timing of the slot-side products (csrc/skinny.cu) against the library fp32 GEMMs torch picks for the same shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from devias_b200 import slot_linear


def timeit(fn, n=50):
    for _ in range(5): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for M, K, N in [(16, 768, 3072), (16, 3072, 768), (16, 768, 2048), (16, 2048, 768), (16, 768, 466), (64, 768, 3072), (64, 3072, 768),
                (64, 768, 2048), (64, 2048, 768), (64, 768, 765), (512, 768, 3072)]:
    x = torch.randn(M, K, device='cuda', requires_grad=True); w = (torch.randn(N, K, device='cuda') * 0.05).requires_grad_(True)
    b = torch.zeros(N, device='cuda', requires_grad=True); dy = torch.randn(M, N, device='cuda')
    row = []
    for name, lin in (('skinny', slot_linear.linear), ('torch', F.linear)):
        with torch.no_grad():
            tf = timeit(lambda: lin(x, w, b))
        def fb():
            y = lin(x, w, b)
            torch.autograd.grad(y, [x, w, b], dy)
        row.append(f'{name}: fwd {tf:6.1f} us  fwd+bwd {timeit(fb):6.1f} us')
    print(f'M={M:3d} K={K:4d} N={N:4d}  ' + '   '.join(row))
H, dh, D = 4, 512, 768
for B, S in [(8, 2), (32, 2)]:
    q = torch.randn(B, S, H, dh, device='cuda', requires_grad=True); wk = (torch.randn(H * dh, D, device='cuda') * 0.05).requires_grad_(True)
    dqt = torch.randn(B, H, S, D, device='cuda')
    def a():
        torch.autograd.grad(slot_linear._FoldKeysFn.apply(q, wk), [q, wk], dqt)
    def t():
        torch.autograd.grad(torch.einsum('bshd,hdc->bhsc', q, wk.view(H, dh, D)), [q, wk], dqt)
    print(f'fold_keys B={B} S={S}: skinny fwd+bwd {timeit(a):6.1f} us   torch {timeit(t):6.1f} us')
