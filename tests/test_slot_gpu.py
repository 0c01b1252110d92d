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
