"""CUDA side of the streaming slot attention (contract: devias_b200/slot_attention.py): csrc/slot_attn.cu forward and
backward.  The gradient w.r.t. the context tokens is summed over the layers of the aggregation block IN PLACE: every
layer's backward kernel accumulates into one shared buffer and only the last one to run hands it to autograd."""
import torch

from . import ops


def token_stats(tokens: torch.Tensor, eps: float = 1e-5):
    """The streaming kernel derives the LayerNorm statistics of the tokens itself; nothing to precompute."""
    return None, None


class _TokenGradSink:
    """shared by the `depth` SlotStreamFn nodes of one AggregationBlock.forward call"""

    def __init__(self):
        self.pending = 0
        self.buf = None


class SlotStreamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, g, G, c0, sink):
        tokens, g, G, c0 = tokens.contiguous(), g.contiguous(), G.contiguous(), c0.contiguous()
        U, m, A, attn, mu, rstd = ops.slot_stream_fwd(tokens, g, G, c0)
        ctx.save_for_backward(tokens, g, G, attn, mu, rstd)
        ctx.c0 = c0
        ctx.sink = sink
        if sink is not None:
            sink.pending += 1
        ctx.mark_non_differentiable(mu, rstd)
        return U, m, A, attn

    @staticmethod
    def backward(ctx, dU, dm, dA, dattn):
        tokens, g, G, attn, mu, rstd = ctx.saved_tensors
        zeros = lambda ref: torch.zeros_like(ref)
        dU = zeros(g) if dU is None else dU
        dm = zeros(G) if dm is None else dm
        dA = zeros(G) if dA is None else dA
        sink = ctx.sink
        need_dt = ctx.needs_input_grad[0]
        if sink is not None and need_dt:
            dt, dg, dG, dc0 = ops.slot_stream_bwd(tokens, mu, rstd, g, G, attn, dU, dm, dA, dattn, dtokens=sink.buf)
            sink.buf = dt
            sink.pending -= 1
            out_dt = None
            if sink.pending == 0:
                out_dt, sink.buf = sink.buf, None
                if tokens.dtype == torch.bfloat16:
                    out_dt = ops.cast_bf16(out_dt)
            return out_dt, dg, dG, dc0, None
        dt, dg, dG, dc0 = ops.slot_stream_bwd(tokens, mu, rstd, g, G, attn, dU, dm, dA, dattn)
        if need_dt and tokens.dtype == torch.bfloat16:
            dt = ops.cast_bf16(dt)
        return (dt if need_dt else None), dg, dG, dc0, None


def new_sink():
    return _TokenGradSink()


def slot_stream(tokens, mu, r, g, G, c0, sink=None):
    """fp32 tokens: the fp32 streaming kernels (1e-5 contract); bf16 tokens: the tcgen05 kernel (bf16 operand rounding)"""
    if tokens.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f'devias_b200 slot attention takes fp32 or bf16 context tokens, got {tokens.dtype}')
    return SlotStreamFn.apply(tokens, g, G, c0, sink)
