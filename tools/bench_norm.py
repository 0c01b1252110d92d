"""HBM-bound helper kernels (LayerNorm forward / backward, bias column sums, weight cast) against the measured copy peak.
Algorithmic bytes = every operand touched once.  GPU box: python tools/bench_norm.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import ops
PEAK = 6549.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')
if os.path.isfile(p):
    PEAK = json.load(open(p)).get('hbm_gbs', PEAK)
def t(fn, n=20):
    if os.environ.get('DEVIAS_ONESHOT'):       # one launch per case for `ncu -c <cases>`
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for B in ([int(a) for a in sys.argv[1:]] or [8, 64]):
    M, D = B * 1568, 768
    x = torch.randn(M, D, device='cuda'); g = torch.randn(D, device='cuda'); b = torch.randn(D, device='cuda')
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-6)
    dy = torch.randn(M, D, device='cuda').bfloat16(); dres = torch.randn(M, D, device='cuda')
    dg = torch.zeros(D, device='cuda'); db = torch.zeros(D, device='cuda'); cs = torch.zeros(D, device='cuda')
    sc = torch.rand(B, device='cuda')
    w32 = torch.randn(86_000_000, device='cuda'); w16 = torch.empty(86_000_000, device='cuda', dtype=torch.bfloat16)
    h = torch.randn(M, 3072, device='cuda').bfloat16(); hb = torch.zeros(3072, device='cuda')
    cases = [
        ('layernorm fwd  (fp32 in, bf16 out, stats)', M * D * (4 + 2) + M * 8, lambda: ops.layernorm_fwd(x, g, b, 1e-6)),
        ('layernorm bwd  (bf16 dy, fp32 x, +resid, fp32 dx in place, scaled bf16 copy, dgamma/dbeta/colsum)', M * D * (2 + 4 + 4 + 4 + 2),
         lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, d_resid=dres, dgamma=dg, dbeta=db, dx_colsum=cs, inplace=True, row_scale=sc, rows_per_scale=1568)),
        ('colsum bf16 [M, 3072]', M * 3072 * 2, lambda: ops.colsum_bf16(h, hb)),
        ('cast fp32 -> bf16, 86 M weights', 86_000_000 * 6, lambda: ops.cast_bf16(w32, w16)),
    ]
    for name, nbytes, fn in cases:
        ms = t(fn)
        print(f'B={B:3d} {name:100s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s ({nbytes / ms / 1e6 / PEAK * 100:4.1f} % of {PEAK:.0f})')
    del w32, w16, h
