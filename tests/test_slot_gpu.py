"""GPU parity of the streaming slot-attention kernel (fp32, <= 1e-5 relative: BASELINE.json north_star) against the
torch evaluation of the same folded contract and -- through AggregationBlock (tests/test_model_gpu.py) -- the reference."""
import pytest
import torch

from oracle import devias_oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('B,N,S', [(1, 1568, 2), (3, 1568, 4), (2, 1568, 8), (2, 100, 2), (5, 16, 2), (1, 1569, 2), (64, 1568, 2)])
def test_slot_stream_fwd(B, N, S):
    from devias_b200 import ops, slot_attention as SA
    HS = 4 * S
    tok = O.synth_tokens(B, n_tokens=N, seed=B + N).cuda() + 0.25
    gen = torch.Generator(device='cuda').manual_seed(S)
    g = torch.randn(B, HS, 768, device='cuda', generator=gen) * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda', generator=gen) * 0.3
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    rmu, rr = SA.token_stats(tok)
    rU, rm, rA, ra = SA.slot_stream_torch(tok.double(), rmu.double(), rr.double(), g.double(), G.double(), c0.double())
    assert_close(mu, rmu, 1e-5, 'mu')
    assert_close(rstd, rr, 1e-5, 'rstd')
    assert_close(attn, ra, 1e-5, 'attn')
    assert_close(A, rA, 1e-5, 'A')
    assert_close(m, rm, 1e-5, 'm')
    assert_close(U, rU, 1e-5, 'U')
    # slot-axis softmax: the S probabilities of every head sum to one for every token
    assert torch.allclose(attn.view(B, 4, S, N).sum(2), torch.ones(B, 4, N, device='cuda'), atol=1e-5)


@pytest.mark.parametrize('B,N,S,with_dattn', [(2, 1568, 2, True), (3, 1568, 4, False), (2, 100, 2, True), (1, 1569, 2, False),
                                               (2, 1568, 8, True), (1, 100, 8, False)])
def test_slot_stream_bwd(B, N, S, with_dattn):
    """streaming backward kernel vs float64 autograd of the same folded contract"""
    from devias_b200 import ops, slot_attention as SA
    HS = 4 * S
    gen = torch.Generator(device='cuda').manual_seed(7 + S)
    tok = O.synth_tokens(B, n_tokens=N, seed=3 * B + N).cuda() + 0.25
    g = torch.randn(B, HS, 768, device='cuda', generator=gen) * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda', generator=gen) * 0.3
    dU = torch.randn(B, HS, 768, device='cuda', generator=gen)
    dm = torch.randn(B, HS, device='cuda', generator=gen)
    dA = torch.randn(B, HS, device='cuda', generator=gen)
    dattn = torch.randn(B, HS, N, device='cuda', generator=gen) if with_dattn else None
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    dt, dg, dG, dc0 = ops.slot_stream_bwd(tok, mu, rstd, g, G, attn, dU, dm, dA, dattn)
    leaves = [t.double().requires_grad_(True) for t in (tok, g, G, c0)]
    rmu, rr = SA.token_stats(leaves[0])
    # token_stats upcasts to float internally for float32 inputs only; recompute in float64 here
    t64 = leaves[0]
    rmu = t64.mean(-1); rr = torch.rsqrt((t64 - rmu.unsqueeze(-1)).square().mean(-1) + 1e-5)
    outs = SA.slot_stream_torch(t64, rmu, rr, leaves[1], leaves[2], leaves[3])
    go = [dU.double(), dm.double(), dA.double(), (dattn.double() if with_dattn else torch.zeros_like(outs[3]))]
    rdt, rdg, rdG, rdc0 = torch.autograd.grad(outs, leaves, go)
    assert_close(dt, rdt, 2e-5, 'd tokens')
    assert_close(dg, rdg, 2e-5, 'dg')
    assert_close(dG, rdG, 2e-5, 'dG')
    assert_close(dc0, rdc0, 2e-5, 'dc0')
    # in-place accumulation across layers
    base = torch.randn_like(tok)
    acc = base.clone()
    ops.slot_stream_bwd(tok, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens=acc)
    assert_close(acc, base.double() + rdt, 2e-5, 'accumulated d tokens')


@pytest.mark.parametrize('B,N,S', [(1, 32, 2), (2, 1568, 2), (3, 100, 4), (1, 1569, 2), (2, 1568, 4), (2, 1568, 8), (3, 100, 8),
                                   (40, 1568, 2)])
def test_slot_stream_fwd_bf16_tokens(B, N, S):
    """BF16 context tokens (BASELINE config 5 'fp32 and bf16'): the tcgen05 kernel of csrc/slot_attn_tc.cu against the float64
    evaluation of the folded contract (agg_block/attention.py:32-40,120-141) on the SAME bf16 token values.  LayerNorm statistics
    are fp32-exact (1e-5); the logits see g rounded to bf16 -- against a reference that rounds g the same way the slot-axis softmax
    agrees to 1e-5, i.e. the tensor-core contraction itself is exact up to fp32 accumulation; against the unrounded reference
    the outputs carry bf16 operand rounding (tolerance 5e-3, inside the 1e-2 bf16 budget of BASELINE.json north_star)."""
    from devias_b200 import ops, slot_attention as SA
    HS = 4 * S
    gen = torch.Generator(device='cuda').manual_seed(11 * S + B)
    tok = (torch.randn(B, N, 768, device='cuda', generator=gen) * (1.0 + torch.rand(B, N, 1, device='cuda', generator=gen))
           + 0.25).to(torch.bfloat16)
    g = torch.randn(B, HS, 768, device='cuda', generator=gen) * 0.05
    G = g.sum(-1).contiguous()
    c0 = torch.randn(B, HS, device='cuda', generator=gen) * 0.3
    U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tok, g, G, c0)
    t64 = tok.double()
    rmu = t64.mean(-1)
    rr = torch.rsqrt((t64 - rmu.unsqueeze(-1)).square().mean(-1) + 1e-5)
    rU, rm, rA, ra = SA.slot_stream_torch(t64, rmu, rr, g.double(), G.double(), c0.double())
    _, _, _, ra_b = SA.slot_stream_torch(t64, rmu, rr, g.to(torch.bfloat16).double(), G.double(), c0.double())
    assert_close(mu, rmu, 1e-5, 'mu')
    assert_close(rstd, rr, 1e-5, 'rstd')
    assert_close(attn, ra_b, 1e-5, 'attn vs reference with bf16-rounded g')
    assert_close(attn, ra, 5e-3, 'attn')
    assert_close(A, rA, 5e-3, 'A')
    assert_close(m, rm, 5e-3, 'm')
    assert_close(U, rU, 5e-3, 'U')
    assert torch.allclose(attn.view(B, 4, S, N).sum(2), torch.ones(B, 4, N, device='cuda'), atol=1e-5)


@pytest.mark.parametrize('B,N,S,with_dattn', [(1, 32, 2, True), (2, 1568, 2, True), (3, 100, 4, False), (1, 1569, 2, False),
                                               (2, 1568, 4, True), (2, 1568, 8, True), (3, 100, 8, False), (40, 1568, 2, False)])
def test_slot_stream_bwd_bf16_tokens(B, N, S, with_dattn):
    """tcgen05 backward for bf16 tokens (csrc/slot_attn_tc_bwd.cu) vs float64 autograd of the folded contract on the same bf16
    token values.  g, dU and the per-token coefficients enter the MMAs as bf16, accumulation and the coefficient math are fp32:
    tolerance 1e-2 (bf16 budget of BASELINE.json north_star; measured 2-3e-3)."""
    from devias_b200 import ops, slot_attention as SA
    HS = 4 * S
    gen = torch.Generator(device='cuda').manual_seed(17 + S + B)
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
    assert dt.dtype == torch.float32
    leaves = [t.double().requires_grad_(True) for t in (tok, g, G, c0)]
    t64 = leaves[0]
    rmu = t64.mean(-1)
    rr = torch.rsqrt((t64 - rmu.unsqueeze(-1)).square().mean(-1) + 1e-5)
    outs = SA.slot_stream_torch(t64, rmu, rr, leaves[1], leaves[2], leaves[3])
    go = [dU.double(), dm.double(), dA.double(), (dattn.double() if with_dattn else torch.zeros_like(outs[3]))]
    rdt, rdg, rdG, rdc0 = torch.autograd.grad(outs, leaves, go)
    assert_close(dt, rdt, 1e-2, 'd tokens')
    assert_close(dg, rdg, 1e-2, 'dg')
    assert_close(dG, rdG, 1e-2, 'dG')
    assert_close(dc0, rdc0, 1e-2, 'dc0')
    base = torch.randn(B, N, 768, device='cuda', generator=gen)
    acc = base.clone()
    ops.slot_stream_bwd(tok, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens=acc)
    assert_close(acc - base, rdt, 1e-2, 'accumulated d tokens')
