"""micro-benchmark of every GEMM shape/epilogue of one encoder block at B clips (GPU box): python tools/bench_gemm.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import ops
from devias_b200.functional import _wgrad_split
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = B * 1568
bf = lambda *s: (torch.randn(*s, device='cuda') * 0.1).bfloat16()
f32 = lambda *s: torch.randn(*s, device='cuda')
def t(fn, n=20):
    if os.environ.get('DEVIAS_ONESHOT'):       # one launch per case: `ncu -c <cases>` then captures every shape exactly once
        fn(); torch.cuda.synchronize(); return 1.0
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x768, x3072, x2304 = bf(M, 768), bf(M, 3072), bf(M, 2304)
wqkv, wproj, w1, w2 = bf(2304, 768), bf(768, 768), bf(3072, 768), bf(768, 3072)
res = f32(M, 768); b768, b2304, b3072 = f32(768), f32(2304), f32(3072)
o_bf = {n: torch.empty(M, n, device='cuda', dtype=torch.bfloat16) for n in (768, 2304, 3072)}
o2 = torch.empty(M, 3072, device='cuda', dtype=torch.bfloat16)
o_f = torch.empty(M, 768, device='cuda')
cases = [
 ('fwd qkv   STORE_BF16 N2304 K768 ', 2*M*2304*768, lambda: ops.gemm(x768, wqkv, ops.EPI_STORE_BF16, bias=b2304, out=o_bf[2304])),
 ('fwd proj  RESID_F32  N768  K768 ', 2*M*768*768, lambda: ops.gemm(x768, wproj, ops.EPI_RESID_F32, bias=b768, aux=res, out=o_f)),
 ('fwd fc1   GELU_BF16  N3072 K768 ', 2*M*3072*768, lambda: ops.gemm(x768, w1, ops.EPI_GELU_BF16, bias=b3072, out=o_bf[3072], out2=o2)),
 ('fwd fc2   RESID_F32  N768  K3072', 2*M*768*3072, lambda: ops.gemm(x3072, w2, ops.EPI_RESID_F32, bias=b768, aux=res, out=o_f)),
 ('dgrad fc2 DGELU_BF16 N3072 K768 ', 2*M*3072*768, lambda: ops.gemm(x768, w2, ops.EPI_DGELU_BF16, b_mn=True, aux=o2, out=o_bf[3072])),
 ('dgrad fc1 STORE_BF16 N768  K3072', 2*M*768*3072, lambda: ops.gemm(x3072, w1, ops.EPI_STORE_BF16, b_mn=True, out=o_bf[768])),
 ('dgrad prj STORE_BF16 N768  K768 ', 2*M*768*768, lambda: ops.gemm(x768, wproj, ops.EPI_STORE_BF16, b_mn=True, out=o_bf[768])),
 ('dgrad qkv STORE_BF16 N768  K2304', 2*M*768*2304, lambda: ops.gemm(x2304, wqkv, ops.EPI_STORE_BF16, b_mn=True, out=o_bf[768])),
]
for (no, ni, xa, xb) in [(768, 3072, x768, x3072), (3072, 768, x3072, x768), (768, 768, x768, x768), (2304, 768, x2304, x768)]:
    g = torch.zeros(no, ni, device='cuda')
    sp = _wgrad_split(no, ni, M)
    cases.append((f'wgrad [{no}x{ni}] split {sp:2d} K{M}', 2*M*no*ni,
                  (lambda xa=xa, xb=xb, g=g, sp=sp: ops.gemm(xa, xb, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=g, split_k=sp))))
tot_t = tot_f = 0
for name, fl, fn in cases:
    ms = t(fn); tot_t += ms; tot_f += fl
    print(f'{name:36s} {ms*1e3:8.1f} us  {fl/ms/1e9:7.0f} TFLOP/s')
print(f'block total {tot_t*1e3:.1f} us -> x12 = {tot_t*12:.2f} ms, avg {tot_f/tot_t/1e9:.0f} TFLOP/s')
