import os, sys
sys.path.insert(0, '/root/repo')
import torch
from devias_b200 import ops
B, S = 64, 2
tok = torch.randn(B, 1568, 768, device='cuda') * 1.5
g = torch.randn(B, 4 * S, 768, device='cuda') * 0.05; G = g.sum(-1).contiguous(); c0 = torch.randn(B, 4 * S, device='cuda')
for _ in range(3): ops.slot_stream_fwd(tok, g, G, c0)
torch.cuda.synchronize()
