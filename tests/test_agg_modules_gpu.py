"""GPU: the aggregation-block MODULES called one by one (the module interface agg_block/attention.py exposes) run on the
hand-written kernels and agree with the oracle's restatement: PreNorm(Attention) (folded streaming layer), a bare Attention
on an un-normalised context (folded slot-row products), FeedForward / PreNorm(FeedForward); any number of slot rows."""
import contextlib
import io

import pytest
import torch

from oracle import devias_oracle as O
from oracle import make_golden as MG
from util import assert_close

pytestmark = pytest.mark.gpu


def _layer(B, S):
    from devias_b200.agg_block.attention import Attention, FeedForward, PreNorm
    sd = {'agg_block.' + k: v for k, v in MG.agg_state(S, 1, True, seed=21).items()}
    p = 'agg_block.layers.0.'
    attn = PreNorm(768, Attention(768, 768, heads=4, dim_head=512), context_dim=768)
    ff = PreNorm(768, FeedForward(768, activation='gelu', mult=4))
    attn.load_state_dict({k[len(p) + 2:]: v for k, v in sd.items() if k.startswith(p + '0.')})
    ff.load_state_dict({k[len(p) + 2:]: v for k, v in sd.items() if k.startswith(p + '2.')})
    return sd, attn.cuda(), ff.cuda()


@pytest.mark.parametrize('B,S', [(2, 2), (40, 2), (3, 8)])
def test_prenorm_modules_on_kernels_vs_oracle(B, S):
    from devias_b200 import _lib
    sd, attn, ff = _layer(B, S)
    x = torch.randn(B, S, 768, generator=torch.Generator().manual_seed(1))
    ctx = O.synth_tokens(B, seed=9)
    n0 = _lib.launch_count()
    out, sim = attn(x.cuda(), context=ctx.cuda(), k_pos=None, q_pos=None)
    y = ff(x.cuda())
    assert _lib.launch_count() - n0 >= 8, 'the modules did not run on the devias_b200 kernels'
    oo, osim = O.slot_cross_attention(sd, 'agg_block.layers.0.0.', x, ctx)
    oy = O.slot_feed_forward(sd, 'agg_block.layers.0.2.', x)
    assert_close(out, oo, 1e-5, 'PreNorm(Attention) out')
    assert_close(sim, osim, 1e-5, 'sim_distill')
    assert_close(y, oy, 1e-5, 'PreNorm(FeedForward)')


def test_bare_attention_on_kernels_matches_torch_algebra_with_gradients():
    from devias_b200 import _lib
    from devias_b200.agg_block.attention import Attention
    torch.manual_seed(3)
    a = Attention(768, 768, heads=4, dim_head=512).cuda()
    x = torch.randn(3, 2, 768, device='cuda', requires_grad=True)
    ctx = (O.synth_tokens(3, n_tokens=300, seed=2).cuda() * 0.3).requires_grad_(True)
    n0 = _lib.launch_count()
    out, sim = a(x, context=ctx)
    assert _lib.launch_count() > n0
    w1, w2 = torch.randn_like(out), torch.randn_like(sim)
    ((out * w1).sum() + (sim * w2).sum()).backward()
    got = [x.grad.clone(), ctx.grad.clone()] + [p.grad.clone() for p in a.parameters()]
    x.grad = ctx.grad = None
    a.zero_grad()
    # the same module in float64 on the plain tensor-algebra branch (double tensors do not take the kernel path)
    a64 = Attention(768, 768, heads=4, dim_head=512).cuda().double()
    a64.load_state_dict({k: v.double() for k, v in a.state_dict().items()})
    x64, c64 = x.detach().double().requires_grad_(True), ctx.detach().double().requires_grad_(True)
    o64, s64 = a64(x64, context=c64)
    ((o64 * w1.double()).sum() + (s64 * w2.double()).sum()).backward()
    want = [x64.grad, c64.grad] + [p.grad for p in a64.parameters()]
    assert_close(out, o64, 1e-5, 'out')
    assert_close(sim, s64, 1e-5, 'sim')
    for g, w, name in zip(got, want, ['dx', 'dcontext', 'dWq', 'dWk', 'dWv', 'dWo', 'dbo']):
        assert_close(g, w, 5e-5, name)
