"""micro-benchmark of the flash attention kernels (GPU box): python tools/bench_attn.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N, H = 1568, 12
qkv = (torch.randn(B * N, 3 * H * 64, device='cuda') * 1.2).bfloat16()
dout = torch.randn(B * N, H * 64, device='cuda').bfloat16()
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
out, lse = ops.flash_attn_fwd(qkv, B, N, H)
fl = 4.0 * B * H * N * N * 64
tf = t(lambda: ops.flash_attn_fwd(qkv, B, N, H))
tb = t(lambda: ops.flash_attn_bwd(qkv, out, dout, lse, B, N, H))
print(f'B={B} fwd {tf*1e3:.1f} us {fl/tf/1e9:.0f} TFLOP/s | bwd {tb*1e3:.1f} us {2.5*fl/tb/1e9:.0f} TFLOP/s  skip_dq={os.environ.get("DEVIAS_DEBUG_SKIP_DQ")}')
