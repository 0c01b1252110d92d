"""Diagnostic for the tcgen05 bf16-token slot-attention kernel (csrc/slot_attn_tc.cu): per-output errors against the float64
evaluation of the folded contract on the SAME bf16 tokens, and timings.  DEVIAS_B200_LIB=<path> loads a -DDV_DEBUG_SPIN build
(a protocol bug then traps instead of hanging the box).  Usage: python tools/check_slot_tc.py [--time]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devias_b200 import _lib, ops, slot_attention as SA  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def case(B, N, S, seed=0):
    HS = 4 * S
    gen = torch.Generator(device='cuda').manual_seed(seed + S)
    tok = (torch.randn(B, N, 768, device='cuda', generator=gen) * (1.0 + torch.rand(B, N, 1, device='cuda', generator=gen))
           + 0.25).to(torch.bfloat16)
    g = torch.randn(B, HS, 768, device='cuda', generator=gen) * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda', generator=gen) * 0.3
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    torch.cuda.synchronize()
    t64 = tok.double()
    rmu = t64.mean(-1)
    rr = torch.rsqrt((t64 - rmu.unsqueeze(-1)).square().mean(-1) + 1e-5)
    rU, rm, rA, ra = SA.slot_stream_torch(t64, rmu, rr, g.double(), G.double(), c0.double())
    # the same with g rounded to bf16 (what the kernel multiplies with): the error floor of phase 1
    gb = g.to(torch.bfloat16).double()
    _, _, _, ra_b = SA.slot_stream_torch(t64, rmu, rr, gb, G.double(), c0.double())
    out = dict(mu=rel(mu, rmu), rstd=rel(rstd, rr), attn=rel(attn, ra), attn_vs_bf16g=rel(attn, ra_b), A=rel(A, rA), m=rel(m, rm),
               U=rel(U, rU), sum1=float((attn.view(B, 4, S, N).sum(2) - 1).abs().max()))
    print(f'B={B} N={N} S={S}: ' + ' '.join(f'{k}={v:.2e}' for k, v in out.items()), flush=True)
    return out


def timing(B, S, iters=20):
    HS = 4 * S
    N = 1568
    tok = torch.randn(B, N, 768, device='cuda').to(torch.bfloat16)
    g = torch.randn(B, HS, 768, device='cuda') * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda') * 0.3
    for _ in range(3):
        ops.slot_stream_fwd(tok, g, G, c0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    U = torch.zeros(B, HS, 768, device='cuda')
    mA = torch.zeros(2, B, HS, device='cuda')
    attn = torch.empty(B, HS, N, device='cuda')
    mu = torch.empty(B, N, device='cuda')
    rs = torch.empty(B, N, device='cuda')
    fn = _lib.lib().devias_slot_stream_fwd_bf16
    st = torch.cuda.current_stream().cuda_stream
    e0.record()
    for _ in range(iters):
        fn(tok.data_ptr(), g.data_ptr(), G.data_ptr(), c0.data_ptr(), U.data_ptr(), mA[0].data_ptr(), mA[1].data_ptr(),
           attn.data_ptr(), mu.data_ptr(), rs.data_ptr(), B, N, 768, S, 1e-5, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    gbs = B * N * 768 * 2 / us / 1e3
    print(f'timing B={B} S={S}: {us:.1f} us  {gbs:.0f} GB/s of bf16 tokens', flush=True)


def case_bwd(B, N, S, with_dattn=True, seed=0):
    HS = 4 * S
    gen = torch.Generator(device='cuda').manual_seed(seed + 3 * S)
    tok = (torch.randn(B, N, 768, device='cuda', generator=gen) * (1.0 + torch.rand(B, N, 1, device='cuda', generator=gen))
           + 0.25).to(torch.bfloat16)
    g = torch.randn(B, HS, 768, device='cuda', generator=gen) * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda', generator=gen) * 0.3
    dU = torch.randn(B, HS, 768, device='cuda', generator=gen)
    dm = torch.randn(B, HS, device='cuda', generator=gen)
    dA = torch.randn(B, HS, device='cuda', generator=gen)
    dattn = torch.randn(B, HS, N, device='cuda', generator=gen) if with_dattn else None
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    dt, dg, dG, dc0 = ops.slot_stream_bwd(tok, mu, rstd, g, G, attn, dU, dm, dA, dattn)
    torch.cuda.synchronize()
    leaves = [t.double().requires_grad_(True) for t in (tok, g, G, c0)]
    t64 = leaves[0]
    rmu = t64.mean(-1)
    rr = torch.rsqrt((t64 - rmu.unsqueeze(-1)).square().mean(-1) + 1e-5)
    outs = SA.slot_stream_torch(t64, rmu, rr, leaves[1], leaves[2], leaves[3])
    go = [dU.double(), dm.double(), dA.double(), (dattn.double() if with_dattn else torch.zeros_like(outs[3]))]
    rdt, rdg, rdG, rdc0 = torch.autograd.grad(outs, leaves, go)
    base = torch.randn(B, N, 768, device='cuda', generator=gen)
    acc = base.clone()
    ops.slot_stream_bwd(tok, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens=acc)
    torch.cuda.synchronize()
    out = dict(dt=rel(dt, rdt), dg=rel(dg, rdg), dG=rel(dG, rdG), dc0=rel(dc0, rdc0), dt_acc=rel(acc - base, rdt))
    print(f'bwd B={B} N={N} S={S} dattn={with_dattn}: ' + ' '.join(f'{k}={v:.2e}' for k, v in out.items()), flush=True)


def timing_bwd(B, S, iters=20):
    HS = 4 * S
    N = 1568
    tok = torch.randn(B, N, 768, device='cuda').to(torch.bfloat16)
    g = torch.randn(B, HS, 768, device='cuda') * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda') * 0.3
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    dU, dm, dA = torch.randn_like(U), torch.randn_like(m), torch.randn_like(A)
    dt = torch.empty(B, N, 768, device='cuda')
    dg = torch.zeros(B, HS, 768, device='cuda')
    dGc = torch.zeros(2, B, HS, device='cuda')
    fn = _lib.lib().devias_slot_stream_bwd_bf16
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: fn(tok.data_ptr(), mu.data_ptr(), rstd.data_ptr(), g.data_ptr(), G.data_ptr(), attn.data_ptr(), dU.data_ptr(),
                      dm.data_ptr(), dA.data_ptr(), None, dt.data_ptr(), 0, dg.data_ptr(), dGc[0].data_ptr(), dGc[1].data_ptr(),
                      B, N, 768, S, st)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    gbs = B * N * 768 * 6 / us / 1e3
    print(f'timing bwd B={B} S={S}: {us:.1f} us  {gbs:.0f} GB/s (bf16 tokens in + fp32 token gradient out)', flush=True)


if __name__ == '__main__':
    if '--bwd' in sys.argv:
        for S in (2, 4, 8):
            for B, N, wd in ((1, 32, True), (2, 1568, True), (3, 100, False), (1, 1569, True), (40, 1568, False)):
                case_bwd(B, N, S, wd)
        if '--time' in sys.argv:
            for S in (2, 4, 8):
                for B in (8, 64, 256):
                    timing_bwd(B, S)
        sys.exit(0)
    for S in (2, 4, 8):
        for B, N in ((1, 32), (2, 1568), (3, 100), (1, 1569), (64, 1568)):
            try:
                case(B, N, S)
            except Exception as e:   # noqa: BLE001
                print(f'B={B} N={N} S={S}: FAILED {type(e).__name__}: {e}', flush=True)
                raise
    if '--time' in sys.argv:
        for S in (2, 4, 8):
            for B in (8, 64, 256):
                timing(B, S)
