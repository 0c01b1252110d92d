"""Slot-attention micro-benchmark (BASELINE.json config 5): 1568 tokens x 768, S in {2,4,8}, B sweep, fp32.
Reports the streaming kernels alone (algorithmic bytes = B*N*768*4 per pass) and the whole
AggregationBlock forward (depth 3, tied) next to the measured HBM peak.  GPU box: python tools/bench_slot.py"""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import ops
from devias_b200.agg_block import AggregationBlock
PEAK = 6549.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.isfile(p):
    PEAK = json.load(open(p)).get('hbm_gbs', PEAK)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N, D = 1568, 768
for S in (2, 4, 8):
    for B in (8, 64, 256):
        HS = 4 * S
        tok = torch.randn(B, N, D, device='cuda') * 1.5
        g = torch.randn(B, HS, D, device='cuda') * 0.05; G = g.sum(-1).contiguous(); c0 = torch.randn(B, HS, device='cuda')
        fwd = t(lambda: ops.slot_stream_fwd(tok, g, G, c0))
        gb = B * N * D * 4 / 1e9
        line = f'S={S} B={B:3d}  stream fwd {fwd*1e3:8.1f} us {gb/fwd*1e3:7.0f} GB/s ({gb/fwd*1e3/PEAK*100:4.1f}% of {PEAK:.0f})'
        if S <= 4:
            U, m, A, attn, mu, r = ops.slot_stream_fwd(tok, g, G, c0)
            dU = torch.randn_like(U); dm = torch.randn_like(m); dA = torch.randn_like(A)
            bwd = t(lambda: ops.slot_stream_bwd(tok, mu, r, g, G, attn, dU, dm, dA))
            line += f' | stream bwd {bwd*1e3:8.1f} us {2*gb/bwd*1e3:7.0f} GB/s'
        with contextlib.redirect_stdout(io.StringIO()):
            m_ = AggregationBlock(num_latents=S, weight_tie_layers=True, depth=3).cuda()
        with torch.no_grad():
            blk = t(lambda: m_(tok), n=5)
        line += f' | AggregationBlock(depth 3) fwd {blk*1e3:8.1f} us = {3*gb/blk*1e3:6.0f} GB/s algorithmic'
        print(line)
        del tok
