"""Encoder self-attention over the packed qkv activation (model/modeling_slot.py:102-112).

INTERIM (round 1): the softmax(QK^T)V core is executed by torch's fused SDPA (a library flash
kernel) while the hand-written tcgen05 flash kernels are brought up; the surrounding qkv / proj
GEMMs are already ours.  The interface below is the one the CUDA kernels will keep.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def attention_fwd(qkv: torch.Tensor, B: int, N: int, H: int, need_grad: bool):
    M, three_d = qkv.shape
    hd = three_d // (3 * H)
    leaf = qkv.detach()
    if need_grad:
        leaf.requires_grad_(True)
    with torch.set_grad_enabled(need_grad):
        v5 = leaf.view(B, N, 3, H, hd)
        q, k, v = (v5[:, :, i].transpose(1, 2) for i in range(3))       # [B, H, N, hd] strided views
        o = F.scaled_dot_product_attention(q, k, v, scale=hd ** -0.5)     # q*scale then softmax(q k^T) v
        out = o.transpose(1, 2).reshape(M, H * hd)
    return out.detach(), (leaf, out) if need_grad else None


def attention_bwd(state, dout: torch.Tensor) -> torch.Tensor:
    leaf, out = state
    (dqkv,) = torch.autograd.grad(out, leaf, dout)
    return dqkv.contiguous()
