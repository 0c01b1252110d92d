"""CUDA side of the streaming slot attention (contract: devias_b200/slot_attention.py).

Forward: csrc/slot_attn.cu (one pass over the tokens per layer).  Backward (round 1): the same folded expressions
differentiated by autograd on the GPU from the saved inputs (INTERIM until the streaming backward kernel lands)."""
import torch

from . import ops
from . import slot_attention as SA


def token_stats(tokens: torch.Tensor, eps: float = 1e-5):
    """The streaming kernel derives the LayerNorm statistics of the tokens itself; nothing to precompute."""
    return None, None


class SlotStreamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, g, G, c0):
        U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tokens.contiguous(), g.contiguous(), G.contiguous(), c0.contiguous())
        ctx.save_for_backward(tokens, g, G, c0)
        return U, m, A, attn

    @staticmethod
    def backward(ctx, dU, dm, dA, dattn):
        tokens, g, G, c0 = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(True) for t in (tokens, g, G, c0)]
            mu, r = SA.token_stats(leaves[0])
            outs = SA.slot_stream_torch(leaves[0], mu, r, *leaves[1:])
            grads = torch.autograd.grad(outs, leaves, [dU, dm, dA, dattn], allow_unused=True)
        return tuple(grads)


def slot_stream(tokens, mu, r, g, G, c0):
    if tokens.dtype != torch.float32:
        tokens = tokens.float()
    return SlotStreamFn.apply(tokens, g, G, c0)
