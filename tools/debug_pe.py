import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200 import _lib
B = 2
clip = torch.ones(B, 3, 16, 224, 224, device='cuda')
W = torch.zeros(768, 1536, device='cuda'); W[5, :] = 1.0; W[300, 7] = 2.0
bias = torch.zeros(768, device='cuda'); pos = torch.zeros(1568, 768, device='cuda')
bias[9] = 3.0; pos[:, 11] = torch.arange(1568, device='cuda').float()
out = torch.full((B * 1568, 768), -7.0, device='cuda')
rc = _lib.lib().devias_patch_embed_fwd(clip.data_ptr(), W.data_ptr(), bias.data_ptr(), pos.data_ptr(), out.data_ptr(), B, 3, 16, 224, 224, 768,
                                       torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print('rc', rc)
o = out.view(B, 1568, 768)
print('unwritten', int((o == -7.0).sum()), 'of', o.numel())
print('col5 unique', torch.unique(o[..., 5])[:10].tolist())
print('col300 unique', torch.unique(o[..., 300])[:10].tolist())
print('col9 unique', torch.unique(o[..., 9])[:10].tolist())
print('col11 first', o[0, :5, 11].tolist(), o[1, 1565:, 11].tolist())
print('col0 unique', torch.unique(o[..., 0])[:10].tolist())
rows_unwritten = (o == -7.0).all(-1)
print('rows fully unwritten', int(rows_unwritten.sum()), rows_unwritten[0].nonzero().flatten()[:20].tolist())
