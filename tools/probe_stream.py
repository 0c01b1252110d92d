"""Ceiling check for the slot kernels: bare TMA token stream vs torch copy / reduction on the same tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import _lib
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, N, D = 256, 1568, 768
tok = torch.randn(B, N, D, device='cuda'); out = torch.empty_like(tok); scratch = torch.zeros(4, device='cuda')
gb = tok.numel() * 4 / 1e9
L = _lib.lib(); st = torch.cuda.current_stream().cuda_stream
for stages in (2, 3, 4):
    ms = t(lambda: _lib.check(L.devias_debug_token_stream(tok.data_ptr(), B, N, stages, scratch.data_ptr(), st), 'probe'))
    print(f'TMA token stream, {stages} stages: {ms*1e3:7.1f} us  {gb/ms*1e3:6.0f} GB/s read')
ms = t(lambda: out.copy_(tok)); print(f'torch copy            : {ms*1e3:7.1f} us  {2*gb/ms*1e3:6.0f} GB/s read+write')
ms = t(lambda: tok.sum());      print(f'torch sum             : {ms*1e3:7.1f} us  {gb/ms*1e3:6.0f} GB/s read')
ms = t(lambda: out.zero_());    print(f'torch zero_           : {ms*1e3:7.1f} us  {gb/ms*1e3:6.0f} GB/s write')
