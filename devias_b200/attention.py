"""Encoder self-attention over the packed qkv activation (model/modeling_slot.py:102-112): hand-written tcgen05
flash-attention kernels (csrc/flash_attn.cu), forward and backward."""
from __future__ import annotations

import torch

from . import ops


def attention_fwd(qkv: torch.Tensor, B: int, N: int, H: int, need_grad: bool):
    out, lse2 = ops.flash_attn_fwd(qkv, B, N, H, need_lse=need_grad)
    return out, ((qkv, out, lse2, B, N, H) if need_grad else None)


def attention_bwd(state, dout: torch.Tensor, delta=None) -> torch.Tensor:
    qkv, out, lse2, B, N, H = state
    return ops.flash_attn_bwd(qkv, out, dout.contiguous(), lse2, B, N, H, delta=delta)
