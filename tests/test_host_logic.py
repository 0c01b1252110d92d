"""Host-side logic that needs no GPU: row maps of the slot-row kernels, the bench contract of the reference arm."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _offsets(view):
    """element offsets of every row (z, m) of a [Z, M.., C] strided view, from torch's own strides"""
    Z = view.shape[0]
    rows = view.reshape(Z, -1, view.shape[-1]) if view.is_contiguous() else None
    out = []
    for z in range(Z):
        sub = view[z]
        idx = torch.cartesian_prod(*[torch.arange(n) for n in sub.shape[:-1]]).reshape(-1, sub.dim() - 1)
        out.append([int(sum(i * s for i, s in zip(ix.tolist(), sub.stride()[:-1]))) + z * view.stride(0) for ix in idx])
    return out


def test_rowmap_matches_torch_strides():
    from devias_b200 import ops
    B, S, H, dh, D = 3, 2, 4, 8, 12
    q = torch.zeros(B, S, H, dh)
    qt = torch.zeros(B, H, S, D)
    x = torch.zeros(5, 16)
    for view in (q.permute(2, 0, 1, 3), qt.permute(1, 0, 2, 3), x.unsqueeze(0), qt.view(B, H * S, D)[:, ::2].permute(1, 0, 2)):
        outer, inner, ld, batch = list(ops.rowmap(view))
        want = _offsets(view)
        for z in range(view.shape[0]):
            M = len(want[z])
            got = [(m // inner) * outer + (m % inner) * ld + z * batch for m in range(M)]
            assert got == want[z], (tuple(view.shape), view.stride())


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores) must print ONE JSON line with the contract keys"""
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['impl'] == 'reference' and d['unit'] == 'clips/s' and d['higher_is_better'] is True and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['metric'].startswith('train clips/sec') and 'workload' in d['config']


def test_agg_block_modules_match_the_oracle_on_cpu():
    """PreNorm(Attention) / PreNorm(FeedForward) called directly (the non-streaming path) against the oracle's restatement of
    agg_block/attention.py, plus cache_fn weight tying.  Plain torch: runs without a GPU."""
    from oracle import devias_oracle as O
    from devias_b200.agg_block.attention import Attention, FeedForward, PreNorm, cache_fn
    sd = O.synth_state_dict(num_latents=2, agg_depth=2, agg_weights_tie=True, depth=0, seed=31)
    p = 'agg_block.layers.0.'
    attn = PreNorm(768, Attention(768, 768, heads=4, dim_head=512), context_dim=768)
    ff = PreNorm(768, FeedForward(768, activation='gelu', mult=4))
    attn.load_state_dict({k[len(p + '0.'):]: v for k, v in sd.items() if k.startswith(p + '0.')})
    ff.load_state_dict({k[len(p + '2.'):]: v for k, v in sd.items() if k.startswith(p + '2.')})
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 2, 768, generator=g)
    ctx = torch.randn(2, 40, 768, generator=g) * 1.5
    with torch.no_grad():
        out, sim = attn(x, context=ctx, k_pos=None, q_pos=None)
        ref_out, ref_sim = O.slot_cross_attention(sd, p + '0.', x, ctx)
        y = ff(x)
        ref_y = O.slot_feed_forward(sd, p + '2.', x)
    assert sim.shape == ref_sim.shape == (2 * 4, 2, 40)
    assert torch.allclose(out, ref_out, rtol=1e-5, atol=1e-5) and torch.allclose(sim, ref_sim, rtol=1e-5, atol=1e-6)
    assert torch.allclose(y, ref_y, rtol=1e-5, atol=1e-5)
    assert torch.allclose(sim.sum(1), torch.ones(8, 40), atol=1e-5)          # the slots compete for every token

    make = cache_fn(lambda: FeedForward(8, activation='relu'))
    a, b, c = make(_cache=True), make(_cache=True), make(_cache=False)
    assert a is b and c is not a
    try:
        FeedForward(8)                                                        # default 'geglu' is rejected, as in the reference
        raise AssertionError('geglu must raise')
    except NotImplementedError:
        pass
