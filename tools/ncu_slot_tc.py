"""ncu target: the tcgen05 bf16-token slot kernel at B = 256 (S from argv, default 2)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from devias_b200 import ops
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B, N, HS = 256, 1568, 4 * S
tok = torch.randn(B, N, 768, device='cuda').to(torch.bfloat16)
g = torch.randn(B, HS, 768, device='cuda') * 0.05
G = g.sum(-1).contiguous()
c0 = torch.randn(B, HS, device='cuda') * 0.3
for _ in range(4):
    ops.slot_stream_fwd(tok, g, G, c0)
torch.cuda.synchronize()
