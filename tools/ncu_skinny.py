"""ncu target: the 64-row slot-side products (csrc/skinny.cu tile_gemm_kernel) of one K400 aggregation layer, forward and backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import slot_linear

M = int(os.environ.get('ROWS', '64'))
for K, N in [(768, 3072), (3072, 768), (768, 2048)]:
    x = torch.randn(M, K, device='cuda', requires_grad=True)
    w = (torch.randn(N, K, device='cuda') * 0.05).requires_grad_(True)
    b = torch.zeros(N, device='cuda', requires_grad=True)
    dy = torch.randn(M, N, device='cuda')
    for _ in range(3):
        y = slot_linear.linear(x, w, b)
        torch.autograd.grad(y, [x, w, b], dy)
torch.cuda.synchronize()
