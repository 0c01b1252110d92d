"""Slot-side products on the CUDA `skinny_*` kernels (csrc/skinny.cu): every projection the aggregation block and the heads
apply to the B*S slot rows -- agg_block/attention.py:120-141 (to_q, to_out, and to_k / to_v folded onto the slots as in
devias_b200/slot_attention.py), :81-82 (FeedForward), model/modeling_slot.py:390-410 (head, mask predictor) -- with their
gradients.  fp32 throughout.  Above MAX_ROWS rows (large evaluation batches) the products are no longer weight-read bound
and go to the library GEMM instead."""
import torch
import torch.nn.functional as F

from . import ops

MAX_ROWS = 32


class _LinearFn(torch.autograd.Function):
    """y = x w^T + b on [M, K] rows"""

    @staticmethod
    def forward(ctx, x, w, b):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        w = w.contiguous()
        y = torch.empty(x2.shape[0], w.shape[0], device=x.device, dtype=torch.float32)
        ops.skinny_nt(x2.unsqueeze(0), w.unsqueeze(0), b, y.unsqueeze(0))
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros_like(x2)
            ops.skinny_nn(dy2.unsqueeze(0), w.unsqueeze(0), dx.unsqueeze(0))
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1] or ctx.has_bias:
            dw, db = ops.skinny_outer(dy2.unsqueeze(0), x2.unsqueeze(0), want_colsum=ctx.has_bias)
            dw = dw[0]
            db = db[0] if ctx.has_bias else None
        return dx, dw, db


def linear(x, w, b=None):
    """F.linear for fp32 slot rows"""
    rows = x.numel() // x.shape[-1]
    if rows > MAX_ROWS or x.dtype != torch.float32 or w.shape[1] % 4 != 0:
        return F.linear(x, w, b)
    return _LinearFn.apply(x, w, b)


class _FoldKeysFn(torch.autograd.Function):
    """qt[b, h, s, :] = Wk_h^T q[b, s, h, :]   (q [B, S, H, dh], wk [H*dh, D])  ->  [B, H, S, D]"""

    @staticmethod
    def forward(ctx, q, wk):
        B, S, H, dh = q.shape
        D = wk.shape[1]
        q, wk = q.contiguous(), wk.contiguous()
        qt = torch.zeros(B, H, S, D, device=q.device, dtype=torch.float32)
        ops.skinny_nn(q.permute(2, 0, 1, 3), wk.view(H, dh, D), qt.permute(1, 0, 2, 3))
        ctx.save_for_backward(q, wk)
        return qt

    @staticmethod
    def backward(ctx, dqt):
        q, wk = ctx.saved_tensors
        B, S, H, dh = q.shape
        D = wk.shape[1]
        dqt = dqt.contiguous()
        dq = torch.empty_like(q)
        ops.skinny_nt(dqt.permute(1, 0, 2, 3), wk.view(H, dh, D), None, dq.permute(2, 0, 1, 3))
        dwk, _ = ops.skinny_outer(q.permute(2, 0, 1, 3), dqt.permute(1, 0, 2, 3))
        return dq, dwk.view(H * dh, D)


class _ApplyValuesFn(torch.autograd.Function):
    """out[b, s, h, :] = Wv_h cbar[b, h, s, :]   (cbar [B, H, S, D], wv [H*dh, D])  ->  [B, S, H*dh]"""

    @staticmethod
    def forward(ctx, cbar, wv):
        B, H, S, D = cbar.shape
        dh = wv.shape[0] // H
        cbar, wv = cbar.contiguous(), wv.contiguous()
        out = torch.empty(B, S, H, dh, device=cbar.device, dtype=torch.float32)
        ops.skinny_nt(cbar.permute(1, 0, 2, 3), wv.view(H, dh, D), None, out.permute(2, 0, 1, 3))
        ctx.save_for_backward(cbar, wv)
        return out.view(B, S, H * dh)

    @staticmethod
    def backward(ctx, dout):
        cbar, wv = ctx.saved_tensors
        B, H, S, D = cbar.shape
        dh = wv.shape[0] // H
        dout = dout.contiguous().view(B, S, H, dh)
        dcbar = torch.zeros_like(cbar)
        ops.skinny_nn(dout.permute(2, 0, 1, 3), wv.view(H, dh, D), dcbar.permute(1, 0, 2, 3))
        dwv, _ = ops.skinny_outer(dout.permute(2, 0, 1, 3), cbar.permute(1, 0, 2, 3))
        return dcbar, dwv.view(H * dh, D)


def fold_keys(q, wk):
    B, S, H, dh = q.shape
    if B * S > MAX_ROWS:
        return torch.einsum('bshd,hdc->bhsc', q, wk.view(H, dh, -1))
    return _FoldKeysFn.apply(q, wk)


def apply_values(cbar, wv):
    B, H, S, D = cbar.shape
    if B * S > MAX_ROWS:
        return torch.einsum('bhsc,hdc->bshd', cbar, wv.view(H, -1, D)).reshape(B, S, -1)
    return _ApplyValuesFn.apply(cbar, wv)
