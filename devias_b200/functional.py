"""torch.autograd.Functions that drive the sm_100a kernels for the encoder of the DEVIAS student
(model/modeling_slot.py:120-152 Block, :155-177 PatchEmbed, :350-377 forward_features).

Every Function launches only kernels from libdevias_b200.so (see include/devias_b200.h).  The
residual stream is fp32, GEMM operands are bf16 (fp32 accumulation in tensor memory).
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch

from . import ops

# --------------------------------------------------------------------------------------------
# direct gradient accumulation: inside `direct_grads()` the backward Functions of this package reduce-add every parameter
# gradient straight into the flat gradient arena (`param._grad_sink`, devias_b200/arena.py) and hand autograd None for it, so
# no per-parameter accumulation kernels, zero fills or temporaries exist.  Used by the captured training step (engine.py),
# where python gradient hooks do not run anyway; eager DDP (hook-driven buckets) keeps the autograd route.
_direct = False


@contextlib.contextmanager
def direct_grads(enable=True):
    global _direct
    prev, _direct = _direct, bool(enable)
    try:
        yield
    finally:
        _direct = prev


def grad_sinks(*params):
    """arena gradient views of `params` if direct accumulation is on and ALL of them live in an arena, else None"""
    if not _direct:
        return None
    out = [getattr(p, '_grad_sink', None) if getattr(p, 'requires_grad', False) else None for p in params]
    return out if all(s is not None for s in out) else None

# --------------------------------------------------------------------------------------------
# side channel: the LayerNorm-backward kernel that produces a residual-stream gradient (fp32) also
# emits its bf16 copy (the operand of the next dgrad/wgrad GEMMs).  autograd only carries the fp32
# tensor between Functions; the copy travels here, tied to the IDENTITY of the producing tensor object (a weak reference --
# a recycled address or a summed / replaced gradient can never match) and its version at the time of the stash.
# The copy may already carry the per-sample drop-path factor of the branch that consumes it next (`scale`, matched by
# identity) and comes with its column sums (= the bias gradient of that branch's last linear layer).
import weakref

_stash = None      # (weakref to dx, version, dxb, scale, colsum)


def _stash_bf16(dx: torch.Tensor, dxb: torch.Tensor, scale=None, colsum=None):
    global _stash
    _stash = (weakref.ref(dx), dx._version, dxb, scale, colsum)


def _match(dx: torch.Tensor):
    t = _stash
    if t is not None and t[0]() is dx and t[1] == dx._version and t[2].shape == dx.shape:
        return t
    return None


def _is_ours(dx: torch.Tensor) -> bool:
    """True when `dx` is the very tensor one of our own backward kernels produced (and stashed): nobody else can alias it."""
    return _match(dx) is not None


def _take_bf16(dx: torch.Tensor) -> torch.Tensor:
    """unscaled bf16 copy of a residual-stream gradient"""
    global _stash
    t = _match(dx)
    _stash = None
    if t is not None and t[3] is None:
        return t[2]
    return ops.cast_bf16(dx.contiguous())


def _take_scaled(dx: torch.Tensor, scale, rows_per_scale: int):
    """(bf16(dx * scale per sample), column sums of it | None): straight from the producing LayerNorm backward when it was told
    this consumer's scale, otherwise by the stand-alone cast kernel."""
    global _stash
    t = _match(dx)
    _stash = None
    if t is not None and t[3] is scale:
        return t[2], t[4]
    if scale is None:
        return ops.cast_bf16(dx.contiguous()), None
    return ops.scale_rows_cast(dx, scale, rows_per_scale), None


def _flat_zeros(like_list, device):
    """one zero-filled fp32 buffer carved into views shaped like `like_list` (8-element aligned)."""
    offs, total = [], 0
    for t in like_list:
        offs.append(total)
        total += (t.numel() + 7) // 8 * 8
    flat = torch.zeros(total, device=device, dtype=torch.float32)
    return [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, like_list)]


def _wgrad_split(rows_out: int, cols_out: int, tokens: int) -> int:
    """split-K factor so that a weight-gradient GEMM (K = tokens) fills the 74 two-CTA clusters about twice."""
    bn = 256 if cols_out % 256 == 0 else 128
    pair_tiles = (((rows_out + 127) // 128 + 1) // 2) * ((cols_out + bn - 1) // bn)
    kblks = (tokens + 63) // 64
    return max(1, min(kblks, (2 * 74 + pair_tiles - 1) // pair_tiles))


# --------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    """Conv3d(k=s=(2,16,16)) + flatten/transpose + sin-cos table add (model/modeling_slot.py:171-177, :354-355)
    as patchify + one tcgen05 GEMM whose epilogue adds bias and the position table."""

    @staticmethod
    def forward(ctx, clip, weight, bias, w16, pos_table):
        B = clip.shape[0]
        D = weight.shape[0]
        implicit = (clip.dtype == torch.float32 and tuple(clip.shape[1:]) == (3, 16, 224, 224) and D == 768
                    and pos_table.dtype == torch.float32 and tuple(pos_table.shape) == (1568, 768))
        if implicit:
            # implicit GEMM: the clip is the A operand (5-D TMA boxes, tf32 MMAs); nothing is materialised or saved but the clip
            clip = clip.contiguous()
            x0 = ops.patch_embed_fwd(clip, weight.detach().contiguous(), bias.detach(), pos_table.contiguous())
            N = x0.shape[0] // B
            ctx.save_for_backward(clip)
        else:
            patches = ops.patchify(clip)                      # [B*N, 1536] bf16
            N = patches.shape[0] // B
            x0 = ops.gemm(patches, w16.view(D, -1), ops.EPI_RESID_F32, bias=bias, aux=pos_table, aux_row_mod=N)
            ctx.save_for_backward(patches)
        ctx.implicit = implicit
        ctx.wshape = weight.shape
        ctx.sinks = grad_sinks(weight, bias)
        return x0.view(B, N, D)

    @staticmethod
    def backward(ctx, dx0):
        (patches,) = ctx.saved_tensors
        if ctx.implicit:
            patches = ops.patchify(patches)                   # the saved tensor is the clip: tube rows for the weight-gradient GEMM
        D = ctx.wshape[0]
        dyb = _take_bf16(dx0).view(-1, D)
        direct = ctx.sinks is not None
        if direct:
            dw, db = ctx.sinks[0].view(D, -1), ctx.sinks[1]
        else:
            dw, db = _flat_zeros([torch.empty(D, patches.shape[1], device='meta'), torch.empty(D, device='meta')], dx0.device)
        ops.gemm(dyb, patches, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=dw,
                 split_k=_wgrad_split(D, patches.shape[1], patches.shape[0]))
        ops.colsum_bf16(dyb, db)
        if direct:
            return None, None, None, None, None
        return None, dw.view(ctx.wshape), db, None, None


# --------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over 768 channels on an fp32 input (final encoder norm, model/modeling_slot.py:373)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype, below_scale=None):
        """below_scale: drop-path factor [B] of the branch that consumes d(x) first (the last block's MLP branch)"""
        x = x.contiguous()
        y, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps, out_dtype)
        ctx.save_for_backward(x, mean, rstd, weight, below_scale)
        ctx.sinks = grad_sinks(weight, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight, below = ctx.saved_tensors
        direct = ctx.sinks is not None
        if direct:
            dg, db = ctx.sinks
            (cs,) = _flat_zeros([weight], x.device)
        else:
            dg, db, cs = _flat_zeros([weight, weight, weight], x.device)
        rows_per = (x.numel() // x.shape[-1]) // below.numel() if below is not None else 1
        dx, dxb = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, weight, dgamma=dg, dbeta=db, dx_colsum=cs, row_scale=below,
                                    rows_per_scale=rows_per)
        _stash_bf16(dx, dxb, below, cs)
        if direct:
            return dx, None, None, None, None, None
        return dx, dg, db, None, None, None


# --------------------------------------------------------------------------------------------
def _attention_fwd(qkv, B, N, H, need_grad):
    """Softmax attention over the packed qkv [B*N, 3*H*hd] bf16 (q scaled by hd^-0.5 BEFORE q.k^T,
    model/modeling_slot.py:105-112) on the tcgen05 flash kernels (csrc/flash_attn.cu).
    Returns (out [B*N, H*hd] bf16, state for backward)."""
    out, lse2 = ops.flash_attn_fwd(qkv, B, N, H, need_lse=need_grad)
    return out, ((qkv, out, lse2, B, N, H) if need_grad else None)


def _attention_bwd(state, dout):
    qkv, out, lse2, B, N, H = state
    return ops.flash_attn_bwd(qkv, out, dout.contiguous(), lse2, B, N, H)


class EncoderBlockFn(torch.autograd.Function):
    """One pre-LN transformer block, forward and backward (model/modeling_slot.py:142-152 with return_attn=False;
    Attention :95-117; Mlp :60-67; DropPath :36-47 as per-sample row scales s1/s2 fused in the residual epilogues)."""

    @staticmethod
    def forward(ctx, x, n1w, n1b, qkv_w, q_bias, v_bias, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b,
                w16, s1, s2, num_heads, eps, below_scale=None):
        """below_scale: the MLP-branch drop-path factor of the block BELOW (the first consumer of this block's d(x)); the last
        LayerNorm backward of this block then hands that block its scaled bf16 operand and fc2-bias gradient ready-made."""
        B, N, D = x.shape
        M = B * N
        x = x.contiguous()
        qkv16, proj16, fc116, fc216 = w16
        need_grad = any(ctx.needs_input_grad)
        xn, mean1, rstd1 = ops.layernorm_fwd(x, n1w, n1b, eps)
        qkv_bias = torch.cat((q_bias, torch.zeros_like(v_bias), v_bias))       # modeling_slot.py:99
        qkv = ops.gemm(xn.view(M, D), qkv16, ops.EPI_STORE_BF16, bias=qkv_bias)
        attn_out, attn_state = _attention_fwd(qkv, B, N, num_heads, need_grad)
        x1 = ops.gemm(attn_out, proj16, ops.EPI_RESID_F32, bias=proj_b, aux=x.view(M, D), row_scale=s1, rows_per_scale=N)
        x1n, mean2, rstd2 = ops.layernorm_fwd(x1, n2w, n2b, eps)
        h_pre, h_act = ops.gemm(x1n, fc116, ops.EPI_GELU_BF16, bias=fc1_b)
        x2 = ops.gemm(h_act, fc216, ops.EPI_RESID_F32, bias=fc2_b, aux=x1, row_scale=s2, rows_per_scale=N)
        if need_grad:
            ctx.save_for_backward(x, mean1, rstd1, xn, attn_out, x1, mean2, rstd2, x1n, h_pre, h_act,
                                  n1w, n2w, qkv16, proj16, fc116, fc216, s1, s2, below_scale)
            ctx.attn_state = attn_state
            ctx.dims = (B, N, D)
            ctx.shapes = [t.shape for t in (n1w, n1b, qkv_w, q_bias, v_bias, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)]
            ctx.sinks = grad_sinks(n1w, n1b, qkv_w, q_bias, v_bias, proj_w, proj_b, n2w, n2b, fc1_w, fc1_b, fc2_w, fc2_b)
        return x2.view(B, N, D)

    @staticmethod
    def backward(ctx, dx2):
        (x, mean1, rstd1, xn, attn_out, x1, mean2, rstd2, x1n, h_pre, h_act,
         n1w, n2w, qkv16, proj16, fc116, fc216, s1, s2, below) = ctx.saved_tensors
        B, N, D = ctx.dims
        M = B * N
        dev = dx2.device
        dx2 = dx2.contiguous()
        ours = _is_ours(dx2)   # our own LN-backward output may be updated in place along the residual chain
        direct = ctx.sinks is not None
        if direct:       # parameter gradients accumulate straight into the gradient arena
            (dn1w, dn1b, dqkv_w, dq_bias, dv_bias, dproj_w, dproj_b, dn2w, dn2b, dfc1_w, dfc1_b, dfc2_w, dfc2_b) = ctx.sinks
            (below_cs,) = _flat_zeros([torch.empty(D, device='meta')], dev)
        else:
            metas = [torch.empty(s, device='meta') for s in ctx.shapes] + [torch.empty(D, device='meta')]
            (dn1w, dn1b, dqkv_w, dq_bias, dv_bias, dproj_w, dproj_b, dn2w, dn2b, dfc1_w, dfc1_b, dfc2_w, dfc2_b, below_cs) = \
                _flat_zeros(metas, dev)
        Hd = fc116.shape[0]
        # ---- MLP branch: x2 = x1 + s2 * (gelu(x1n W1^T + b1) W2^T + b2)
        dyb, cs = _take_scaled(dx2, s2, N)
        dyb = dyb.view(M, D)
        dh = ops.gemm(dyb, fc216, ops.EPI_DGELU_BF16, b_mn=True, aux=h_pre)            # [M, Hd]
        ops.gemm(dyb, h_act, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=dfc2_w, split_k=_wgrad_split(D, Hd, M))
        if cs is not None and direct:
            dfc2_b.add_(cs.view(dfc2_b.shape))
        elif cs is not None:
            dfc2_b = cs.view(dfc2_b.shape)           # column sums came with the operand from the producing LayerNorm backward
        else:
            ops.colsum_bf16(dyb, dfc2_b)
        dx1n = ops.gemm(dh, fc116, ops.EPI_STORE_BF16, b_mn=True)                        # [M, D]
        ops.gemm(dh, x1n, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=dfc1_w, split_k=_wgrad_split(Hd, D, M))
        ops.colsum_bf16(dh, dfc1_b)
        del dh
        # (its bf16 copy carries this block's attention-branch factor s1, its column sums are d proj.bias)
        dx1, dyb = ops.layernorm_bwd(dx1n, x1, mean2, rstd2, n2w, d_resid=dx2.view(M, D), dgamma=dn2w, dbeta=dn2b,
                                      dx_colsum=dproj_b, inplace=ours, row_scale=s1, rows_per_scale=N)
        # ---- attention branch: x1 = x + s1 * (attn(xn) Wp^T + bp)
        dattn = ops.gemm(dyb, proj16, ops.EPI_STORE_BF16, b_mn=True)                     # [M, D]
        ops.gemm(dyb, attn_out, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=dproj_w, split_k=_wgrad_split(D, D, M))
        dqkv = _attention_bwd(ctx.attn_state, dattn)                                     # [M, 3D] bf16
        ctx.attn_state = None
        dxn = ops.gemm(dqkv, qkv16, ops.EPI_STORE_BF16, b_mn=True)                       # [M, D]
        ops.gemm(dqkv, xn.view(M, D), ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=dqkv_w, split_k=_wgrad_split(3 * D, D, M))
        ops.colsum_bf16(dqkv[:, :D], dq_bias)
        ops.colsum_bf16(dqkv[:, 2 * D:], dv_bias)
        dx, dxb = ops.layernorm_bwd(dxn, x.view(M, D), mean1, rstd1, n1w, d_resid=dx1, dgamma=dn1w, dbeta=dn1b,
                                    dx_colsum=below_cs, inplace=True, row_scale=below, rows_per_scale=N)
        dx = dx.view(B, N, D)
        _stash_bf16(dx, dxb.view(B, N, D), below, below_cs)
        if direct:
            return (dx,) + (None,) * 19
        return (dx, dn1w, dn1b, dqkv_w, dq_bias, dv_bias, dproj_w, dproj_b, dn2w, dn2b, dfc1_w, dfc1_b,
                dfc2_w, dfc2_b, None, None, None, None, None, None)
