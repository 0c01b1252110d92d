"""One forward and one backward streaming pass (B = 256; S and the token dtype from the environment: S=2|4|8, TOKENS=f32|bf16)
for an ncu capture:
ncu --set full --clock-control none --import-source on -k regex:slot_stream -s 2 -c 2 -o gpurun_out/slot python tools/ncu_slot.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import ops
B, S, N, D = int(os.environ.get('B', 256)), int(os.environ.get('S', 2)), 1568, 768
HS = 4 * S
tok = (torch.randn(B, N, D, device='cuda') * 1.5).to(torch.bfloat16 if os.environ.get('TOKENS', 'f32') == 'bf16' else torch.float32)
g = torch.randn(B, HS, D, device='cuda') * 0.05; G = g.sum(-1).contiguous(); c0 = torch.randn(B, HS, device='cuda')
for _ in range(2):
    U, m, A, attn, mu, r = ops.slot_stream_fwd(tok, g, G, c0)
    dU = torch.randn_like(U); dm = torch.randn_like(m); dA = torch.randn_like(A)
    ops.slot_stream_bwd(tok, mu, r, g, G, attn, dU, dm, dA)
torch.cuda.synchronize()
