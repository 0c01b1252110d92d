"""Slot-side products on the CUDA `skinny_*` kernels (csrc/skinny.cu): every projection the aggregation block and the heads
apply to the B*S slot rows -- agg_block/attention.py:120-141 (to_q, to_out, and to_k / to_v folded onto the slots as in
devias_b200/slot_attention.py), :81-82 (FeedForward), model/modeling_slot.py:390-410 (head, mask predictor) -- with their
gradients.  fp32 throughout, any number of rows (the kernels tile over the rows; weights are re-read from L2 per 16-row
tile).  Inside `functional.direct_grads()` weight / bias gradients are accumulated straight into the gradient arena (which is
also how the contributions of weight-tied layers are summed without extra launches)."""
import torch

from . import ops
from .functional import grad_sinks


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
        ctx.sinks = grad_sinks(w, b) if b is not None else grad_sinks(w)
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
        if ctx.sinks is not None:
            ops.skinny_outer(dy2.unsqueeze(0), x2.unsqueeze(0), into=ctx.sinks[0], colsum_into=ctx.sinks[1] if ctx.has_bias else None)
        elif ctx.needs_input_grad[1] or ctx.has_bias:
            dw, db = ops.skinny_outer(dy2.unsqueeze(0), x2.unsqueeze(0), want_colsum=ctx.has_bias)
            dw = dw[0]
            db = db[0] if ctx.has_bias else None
        return dx, dw, db


def linear(x, w, b=None):
    """F.linear for fp32 slot rows"""
    if not x.is_cuda:
        raise RuntimeError('devias_b200.slot_linear runs on CUDA only (no CPU fallback)')
    if x.dtype != torch.float32:
        x = x.float()
    assert w.shape[1] % 4 == 0, 'slot-row products need an input width that is a multiple of 4'
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
        ctx.sinks = grad_sinks(wk)
        return qt

    @staticmethod
    def backward(ctx, dqt):
        q, wk = ctx.saved_tensors
        B, S, H, dh = q.shape
        D = wk.shape[1]
        dqt = dqt.contiguous()
        dq = torch.empty_like(q)
        ops.skinny_nt(dqt.permute(1, 0, 2, 3), wk.view(H, dh, D), None, dq.permute(2, 0, 1, 3))
        if ctx.sinks is not None:
            ops.skinny_outer(q.permute(2, 0, 1, 3), dqt.permute(1, 0, 2, 3), into=ctx.sinks[0])
            return dq, None
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
        ctx.sinks = grad_sinks(wv)
        return out.view(B, S, H * dh)

    @staticmethod
    def backward(ctx, dout):
        cbar, wv = ctx.saved_tensors
        B, H, S, D = cbar.shape
        dh = wv.shape[0] // H
        dout = dout.contiguous().view(B, S, H, dh)
        dcbar = torch.zeros_like(cbar)
        ops.skinny_nn(dout.permute(2, 0, 1, 3), wv.view(H, dh, D), dcbar.permute(1, 0, 2, 3))
        if ctx.sinks is not None:
            ops.skinny_outer(dout.permute(2, 0, 1, 3), cbar.permute(1, 0, 2, 3), into=ctx.sinks[0])
            return dcbar, None
        dwv, _ = ops.skinny_outer(dout.permute(2, 0, 1, 3), cbar.permute(1, 0, 2, 3))
        return dcbar, dwv.view(H * dh, D)


def fold_keys(q, wk):
    return _FoldKeysFn.apply(q, wk)


def apply_values(cbar, wv):
    return _ApplyValuesFn.apply(cbar, wv)


# --------------------------------------------------------------------------------------------
# the O(S) algebra around the streaming kernel (csrc/slot_glue.cu): one launch per direction instead of ~15 elementwise /
# reduction launches each
class _FoldEpilogueFn(torch.autograd.Function):
    """(qt [B,H,S,D], gamma, beta, scale) -> g [B,H*S,D] = scale qt gamma, G [B,H*S] = sum_c g, c0 [B,H*S] = scale qt . beta"""

    @staticmethod
    def forward(ctx, qt, gamma, beta, scale):
        B, H, S, D = qt.shape
        qt, gamma, beta = qt.contiguous(), gamma.contiguous(), beta.contiguous()
        R = B * H * S
        g = torch.empty(B, H * S, D, device=qt.device, dtype=torch.float32)
        G = torch.empty(B, H * S, device=qt.device, dtype=torch.float32)
        c0 = torch.empty(B, H * S, device=qt.device, dtype=torch.float32)
        ops.check(ops.L().devias_slot_fold_fwd(qt.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(scale), g.data_ptr(),
                                               G.data_ptr(), c0.data_ptr(), R, D, ops.stream()), 'slot_fold_fwd')
        ctx.save_for_backward(qt, gamma, beta)
        ctx.scale = float(scale)
        ctx.sinks = grad_sinks(gamma, beta)
        return g, G, c0

    @staticmethod
    def backward(ctx, dg, dG, dc0):
        qt, gamma, beta = ctx.saved_tensors
        B, H, S, D = qt.shape
        R = B * H * S
        z = lambda ref_shape: torch.zeros(ref_shape, device=qt.device, dtype=torch.float32)
        dg = z((B, H * S, D)) if dg is None else dg.contiguous()
        dG = z((B, H * S)) if dG is None else dG.contiguous()
        dc0 = z((B, H * S)) if dc0 is None else dc0.contiguous()
        dqt = torch.empty_like(qt)
        direct = ctx.sinks is not None
        dgb = ctx.sinks if direct else torch.zeros(2, D, device=qt.device, dtype=torch.float32)
        ops.check(ops.L().devias_slot_fold_bwd(qt.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ctx.scale, dg.data_ptr(), dG.data_ptr(),
                                               dc0.data_ptr(), dqt.data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), R, D, ops.stream()),
                  'slot_fold_bwd')
        if direct:
            return dqt, None, None, None
        return dqt, dgb[0], dgb[1], None


class _ContextFn(torch.autograd.Function):
    """cbar = (gamma (U - m) + beta A) / (A + eps);  U [B,HS,D], m, A [B,HS]"""

    @staticmethod
    def forward(ctx, U, m, A, gamma, beta, eps):
        U, m, A, gamma, beta = U.contiguous(), m.contiguous(), A.contiguous(), gamma.contiguous(), beta.contiguous()
        B, HS, D = U.shape
        cbar = torch.empty_like(U)
        ops.check(ops.L().devias_slot_ctx_fwd(U.data_ptr(), m.data_ptr(), A.data_ptr(), gamma.data_ptr(), beta.data_ptr(), float(eps),
                                              cbar.data_ptr(), B * HS, D, ops.stream()), 'slot_ctx_fwd')
        ctx.save_for_backward(U, m, A, gamma, beta)
        ctx.eps = float(eps)
        ctx.sinks = grad_sinks(gamma, beta)
        return cbar

    @staticmethod
    def backward(ctx, dcbar):
        U, m, A, gamma, beta = ctx.saved_tensors
        B, HS, D = U.shape
        dcbar = dcbar.contiguous()
        dU = torch.empty_like(U)
        dmA = torch.empty(2, B, HS, device=U.device, dtype=torch.float32)
        direct = ctx.sinks is not None
        dgb = ctx.sinks if direct else torch.zeros(2, D, device=U.device, dtype=torch.float32)
        ops.check(ops.L().devias_slot_ctx_bwd(dcbar.data_ptr(), U.data_ptr(), m.data_ptr(), A.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                              ctx.eps, dU.data_ptr(), dmA[0].data_ptr(), dmA[1].data_ptr(), dgb[0].data_ptr(),
                                              dgb[1].data_ptr(), B * HS, D, ops.stream()), 'slot_ctx_bwd')
        if direct:
            return dU, dmA[0], dmA[1], None, None, None
        return dU, dmA[0], dmA[1], dgb[0], dgb[1], None


def fold_epilogue(qt, gamma, beta, scale):
    return _FoldEpilogueFn.apply(qt, gamma, beta, scale)


def context(U, m, A, gamma, beta, eps=1e-7):
    return _ContextFn.apply(U, m, A, gamma, beta, eps)


class _SlotLayerNormFn(torch.autograd.Function):
    """nn.LayerNorm on fp32 slot rows (agg_block/attention.py:29-30,35; agg_block/agg_block.py:111) through the LayerNorm kernels
    of csrc/layernorm.cu instead of the ATen ones; gamma / beta gradients accumulate in place."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps, torch.float32)
        ctx.save_for_backward(x, mean, rstd, weight)
        ctx.sinks = grad_sinks(weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight = ctx.saved_tensors
        direct = ctx.sinks is not None
        dgb = ctx.sinks if direct else torch.zeros(2, weight.numel(), device=x.device, dtype=torch.float32)
        dx, _ = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, weight, want_bf16=False, dgamma=dgb[0], dbeta=dgb[1])
        if direct:
            return dx, None, None, None
        return dx, dgb[0], dgb[1], None


def layer_norm(x, norm: torch.nn.LayerNorm):
    """`norm(x)` for fp32 CUDA slot rows of width 768"""
    if x.shape[-1] != 768 or not norm.elementwise_affine:
        raise NotImplementedError('the LayerNorm kernels are instantiated for width 768 with affine parameters (DEVIAS slot dim)')
    return _SlotLayerNormFn.apply(x.float(), norm.weight, norm.bias, norm.eps)


class _BmmNTFn(torch.autograd.Function):
    """y[z] = x[z] w[z]^T   (x [Z, M, K], w [Z, N, K]) -> [Z, M, N]; both operands may need gradients"""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        y = torch.empty(x.shape[0], x.shape[1], w.shape[1], device=x.device, dtype=torch.float32)
        ops.skinny_nt(x, w, None, y)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.zeros_like(x)
            ops.skinny_nn(dy, w, dx)
        if ctx.needs_input_grad[1]:
            dw, _ = ops.skinny_outer(dy, x)
        return dx, dw


class _BmmNNFn(torch.autograd.Function):
    """y[z] = x[z] w[z]   (x [Z, M, K], w [Z, K, N]) -> [Z, M, N]"""

    @staticmethod
    def forward(ctx, x, w):
        x, w = x.contiguous(), w.contiguous()
        y = torch.zeros(x.shape[0], x.shape[1], w.shape[2], device=x.device, dtype=torch.float32)
        ops.skinny_nn(x, w, y)
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            ops.skinny_nt(dy, w, None, dx)
        if ctx.needs_input_grad[1]:
            dw, _ = ops.skinny_outer(x, dy)
        return dx, dw


def bmm_nt(x, w):
    return _BmmNTFn.apply(x, w)


def bmm_nn(x, w):
    return _BmmNNFn.apply(x, w)
