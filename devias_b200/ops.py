"""Raw (non-autograd) tensor-level wrappers over the C-ABI.  Every function launches hand-written
sm_100a kernels on the current torch CUDA stream; none has a CPU or PyTorch fallback."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

EPI_STORE_BF16, EPI_STORE_F32, EPI_GELU_BF16, EPI_DGELU_BF16, EPI_RESID_F32, EPI_ATOMIC_F32 = range(6)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# small public aliases for modules that bind further entry points themselves (slot_linear.py)
stream = _stream
check = _lib.check


def L():
    return _lib.lib()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('devias_b200 ops need CUDA tensors (there is no CPU fallback)')


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int, *, a_mn=False, b_mn=False, out=None, out2=None, bias=None,
         aux=None, aux_row_mod=0, row_scale=None, rows_per_scale=0, split_k=1):
    """D[m,n] = sum_k A[m,k] B[n,k] with a fused epilogue (include/devias_b200.h: devias_gemm_bf16).
    a: [M,K] (a_mn=False) or [K,M] (a_mn=True); b: [N,K] or [K,N]; 2-D, unit inner stride, bf16."""
    _need_cuda(a, b, out, out2, bias, aux, row_scale)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.dim() == 2 and b.dim() == 2
    assert a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, (a.shape, b.shape, a_mn, b_mn)
    f32_out = epilogue in (EPI_STORE_F32, EPI_RESID_F32, EPI_ATOMIC_F32)
    if out is None:
        if epilogue == EPI_ATOMIC_F32:
            out = torch.zeros(M, N, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    assert out.shape == (M, N) and out.stride(1) == 1 and out.dtype == (torch.float32 if f32_out else torch.bfloat16)
    if epilogue == EPI_GELU_BF16 and out2 is None:
        out2 = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
    if aux is not None:
        assert aux.stride(1) == 1
        assert aux.dtype == (torch.float32 if epilogue == EPI_RESID_F32 else torch.bfloat16)
    if row_scale is not None:
        assert row_scale.dtype == torch.float32 and row_scale.is_contiguous()
    rc = _lib.lib().devias_gemm_bf16(
        a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), M, N, K, epilogue,
        out.data_ptr(), out.stride(0), _ptr(out2), 0 if out2 is None else out2.stride(0), _ptr(bias),
        _ptr(aux), 0 if aux is None else aux.stride(0), aux_row_mod, _ptr(row_scale), rows_per_scale, split_k, _stream())
    _lib.check(rc, 'gemm_bf16')
    return (out, out2) if epilogue == EPI_GELU_BF16 else out


def layernorm_fwd(x: torch.Tensor, gamma, beta, eps: float, out_dtype=torch.bfloat16, want_stats=True):
    """x fp32 [..., 768] -> (y, mean, rstd) (include/devias_b200.h: devias_layernorm_fwd)."""
    _need_cuda(x, gamma, beta)
    assert x.dtype == torch.float32 and x.is_contiguous()
    D = x.shape[-1]
    rows = x.numel() // D
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    assert out_dtype in (torch.bfloat16, torch.float32)
    rc = _lib.lib().devias_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                         int(out_dtype == torch.bfloat16), _ptr(mean), _ptr(rstd), rows, D, float(eps), _stream())
    _lib.check(rc, 'layernorm_fwd')
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, d_resid=None, want_dx=True, want_bf16=True, dgamma=None, dbeta=None,
                  dx_colsum=None, inplace=False, row_scale=None, rows_per_scale=1):
    """returns (dx fp32 | None, dx bf16 | None); dgamma/dbeta/dx_colsum are accumulated in place when given.
    inplace=True writes dx over d_resid (only for gradient buffers this package produced itself).
    row_scale [rows / rows_per_scale] multiplies the bf16 copy and dx_colsum (the consumer's drop-path factor), not dx."""
    _need_cuda(dy, x)
    assert dy.is_contiguous() and x.is_contiguous() and dy.dtype in (torch.bfloat16, torch.float32)
    D = x.shape[-1]
    rows = x.numel() // D
    dx = None
    if want_dx:
        dx = d_resid if (inplace and d_resid is not None) else torch.empty_like(x)
    if d_resid is not None:
        assert d_resid.dtype == torch.float32 and d_resid.is_contiguous()
    dxb = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    rc = _lib.lib().devias_layernorm_bwd(dy.data_ptr(), int(dy.dtype == torch.bfloat16), x.data_ptr(), mean.data_ptr(),
                                         rstd.data_ptr(), gamma.data_ptr(), _ptr(d_resid), _ptr(dx), _ptr(dxb), _ptr(dgamma),
                                         _ptr(dbeta), _ptr(dx_colsum), _ptr(row_scale), int(rows_per_scale), rows, D, _stream())
    _lib.check(rc, 'layernorm_bwd')
    return dx, dxb


def colsum_bf16(a: torch.Tensor, out: torch.Tensor):
    """out[c] += sum_r a[r, c]; a bf16 2-D (row stride allowed), out fp32 [cols]."""
    _need_cuda(a, out)
    assert a.dtype == torch.bfloat16 and a.dim() == 2 and a.stride(1) == 1 and out.dtype == torch.float32
    rc = _lib.lib().devias_colsum_bf16(a.data_ptr(), a.stride(0), a.shape[0], a.shape[1], out.data_ptr(), _stream())
    _lib.check(rc, 'colsum_bf16')
    return out


def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None):
    _need_cuda(src, dst)
    assert src.dtype == torch.float32 and src.is_contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    assert dst.is_contiguous() and dst.numel() == src.numel()
    rc = _lib.lib().devias_cast_f32_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
    _lib.check(rc, 'cast_f32_bf16')
    return dst


def cast_f32(src: torch.Tensor):
    """bf16 -> fp32 copy"""
    _need_cuda(src)
    assert src.dtype == torch.bfloat16 and src.is_contiguous()
    dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    rc = _lib.lib().devias_cast_bf16_f32(src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
    _lib.check(rc, 'cast_bf16_f32')
    return dst


def scale_rows_cast(src: torch.Tensor, row_scale: torch.Tensor, rows_per_scale: int):
    _need_cuda(src, row_scale)
    assert src.dtype == torch.float32 and src.is_contiguous()
    cols = src.shape[-1]
    rows = src.numel() // cols
    dst = torch.empty(src.shape, device=src.device, dtype=torch.bfloat16)
    rc = _lib.lib().devias_scale_rows_cast(src.data_ptr(), dst.data_ptr(), rows, cols, row_scale.data_ptr(), rows_per_scale, _stream())
    _lib.check(rc, 'scale_rows_cast')
    return dst


_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def patchify(clip: torch.Tensor):
    """clip [B,C,T,H,W] (f32/bf16/f16) -> bf16 [B*T/2*H/16*W/16, C*512]"""
    _need_cuda(clip)
    assert clip.dim() == 5 and clip.dtype in _DT
    clip = clip.contiguous()
    B, C, T, H, W = clip.shape
    out = torch.empty(B * (T // 2) * (H // 16) * (W // 16), C * 512, device=clip.device, dtype=torch.bfloat16)
    rc = _lib.lib().devias_patchify(clip.data_ptr(), _DT[clip.dtype], out.data_ptr(), B, C, T, H, W, _stream())
    _lib.check(rc, 'patchify')
    return out


def patch_embed_fwd(clip: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, pos: torch.Tensor):
    """clip fp32 [B,3,16,224,224], weight fp32 [768,3,2,16,16], bias [768], pos fp32 [1568,768] -> fp32 [B*1568, 768]
    (include/devias_b200.h: devias_patch_embed_fwd -- implicit GEMM, the clip is read in place by TMA)"""
    _need_cuda(clip, weight, bias, pos)
    assert clip.dtype == weight.dtype == pos.dtype == torch.float32 and clip.is_contiguous() and weight.is_contiguous() and pos.is_contiguous()
    B, C, T, H, W = clip.shape
    D = weight.shape[0]
    out = torch.empty(B * (T // 2) * (H // 16) * (W // 16), D, device=clip.device, dtype=torch.float32)
    rc = _lib.lib().devias_patch_embed_fwd(clip.data_ptr(), weight.data_ptr(), bias.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                           B, C, T, H, W, D, _stream())
    _lib.check(rc, 'patch_embed_fwd')
    return out


def _rowmap(outer, inner, ld, batch):
    import ctypes
    return (ctypes.c_int64 * 4)(int(outer), int(inner), int(ld), int(batch))


def rowmap(view: torch.Tensor):
    """Row map of a [Z, M, C] or [Z, M1, M2, C] strided VIEW (unit stride on C): problem z, row m = m1 * M2 + m2."""
    assert view.stride(-1) == 1 or view.shape[-1] == 1
    if view.dim() == 3:
        return _rowmap(0, max(view.shape[1], 1), view.stride(1), view.stride(0))
    assert view.dim() == 4
    return _rowmap(view.stride(1), view.shape[2], view.stride(2), view.stride(0))


def skinny_nt(x: torch.Tensor, w: torch.Tensor, bias, y: torch.Tensor):
    """y[z, m, :] = x[z, m, :] @ w[z].T (+ bias);  x, y: fp32 strided views [Z, M.., K] / [Z, M.., N]; w [Z, N, K] contiguous"""
    _need_cuda(x)
    Z, N, K = w.shape
    M = x.numel() // (Z * K)
    assert x.dtype == w.dtype == y.dtype == torch.float32 and w.is_contiguous() and x.shape[-1] == K and y.shape[-1] == N
    assert y.numel() == Z * M * N and x.shape[0] == Z and y.shape[0] == Z
    rc = _lib.lib().devias_skinny_nt(x.data_ptr(), rowmap(x), w.data_ptr(), N * K, _ptr(bias), y.data_ptr(), rowmap(y), M, N, K, Z,
                                     _stream())
    _lib.check(rc, 'skinny_nt')
    return y


def skinny_nn(x: torch.Tensor, w: torch.Tensor, y: torch.Tensor):
    """y[z, m, :] += x[z, m, :] @ w[z];  w [Z, K, N] contiguous; y must be pre-filled"""
    _need_cuda(x)
    Z, K, N = w.shape
    M = x.numel() // (Z * K)
    assert x.dtype == w.dtype == y.dtype == torch.float32 and w.is_contiguous() and x.shape[-1] == K and y.shape[-1] == N
    assert y.numel() == Z * M * N and x.shape[0] == Z and y.shape[0] == Z
    rc = _lib.lib().devias_skinny_nn(x.data_ptr(), rowmap(x), w.data_ptr(), K * N, y.data_ptr(), rowmap(y), M, N, K, Z, _stream())
    _lib.check(rc, 'skinny_nn')
    return y


def skinny_outer(a: torch.Tensor, b: torch.Tensor, want_colsum=False, into=None, colsum_into=None):
    """c[z] = a[z].T @ b[z]  ([Z, I, J]);  optionally colsum[z, i] = sum_m a[z, m, i].
    into / colsum_into: contiguous [Z, I, J] / [Z, I] buffers that are ACCUMULATED into (+=) instead of allocating results."""
    _need_cuda(a)
    Z, I, J = a.shape[0], a.shape[-1], b.shape[-1]
    M = a.numel() // (Z * I)
    assert a.dtype == b.dtype == torch.float32 and b.shape[0] == Z and b.numel() == Z * M * J
    acc = into is not None
    if acc:
        assert into.is_contiguous() and into.numel() == Z * I * J and into.dtype == torch.float32
        c, cs = into, colsum_into
        assert cs is None or (cs.is_contiguous() and cs.numel() == Z * I)
    else:
        c = torch.empty(Z, I, J, device=a.device, dtype=torch.float32)
        cs = torch.empty(Z, I, device=a.device, dtype=torch.float32) if want_colsum else None
    rc = _lib.lib().devias_skinny_outer(a.data_ptr(), rowmap(a), b.data_ptr(), rowmap(b), c.data_ptr(), I * J, _ptr(cs), I, M, I, J, Z,
                                        int(acc), _stream())
    _lib.check(rc, 'skinny_outer')
    return c, cs


def slot_select(slots_head: torch.Tensor, num_slots: int, n_action: int, n_scene: int):
    """slots_head fp32 [B*S, >= n_action + n_scene] -> (action slot index [B], scene slot index [B]) int64"""
    _need_cuda(slots_head)
    assert slots_head.dtype == torch.float32 and slots_head.dim() == 2 and slots_head.stride(1) == 1
    B = slots_head.shape[0] // num_slots
    a = torch.empty(B, device=slots_head.device, dtype=torch.int64)
    s = torch.empty(B, device=slots_head.device, dtype=torch.int64)
    rc = _lib.lib().devias_slot_select(slots_head.data_ptr(), slots_head.stride(0), B, num_slots, n_action, n_scene, a.data_ptr(),
                                       s.data_ptr(), _stream())
    _lib.check(rc, 'slot_select')
    return a, s


def flash_attn_fwd(qkv: torch.Tensor, B: int, N: int, H: int, need_lse=True):
    """qkv bf16 [B*N, 3*H*64] -> (out bf16 [B*N, H*64], lse2 fp32 [B, H, Npad] | None)"""
    _need_cuda(qkv)
    assert qkv.dtype == torch.bfloat16 and qkv.is_contiguous() and qkv.shape == (B * N, 3 * H * 64)
    out = torch.empty(B * N, H * 64, device=qkv.device, dtype=torch.bfloat16)
    npad = (N + 127) // 128 * 128
    lse2 = torch.empty(B, H, npad, device=qkv.device, dtype=torch.float32) if need_lse else None
    rc = _lib.lib().devias_flash_attn_fwd(qkv.data_ptr(), out.data_ptr(), _ptr(lse2), B, N, H, 64, 0.125, _stream())
    _lib.check(rc, 'flash_attn_fwd')
    return out, lse2


def flash_attn_bwd(qkv, out, dout, lse2, B: int, N: int, H: int):
    """-> dqkv bf16 [B*N, 3*H*64]  (include/devias_b200.h: devias_flash_attn_bwd)"""
    _need_cuda(qkv, out, dout, lse2)
    assert dout.dtype == torch.bfloat16 and dout.is_contiguous() and out.is_contiguous() and qkv.is_contiguous()
    npad = (N + 127) // 128 * 128
    dqkv = torch.empty_like(qkv)
    aug = torch.empty(B * H * npad * 16, device=qkv.device, dtype=torch.bfloat16)     # [lse | delta] k-step operand blocks
    dq = torch.empty(B * N * H * 64, device=qkv.device, dtype=torch.float32)
    rc = _lib.lib().devias_flash_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse2.data_ptr(), dqkv.data_ptr(),
                                          aug.data_ptr(), dq.data_ptr(), B, N, H, 64, 0.125, _stream())
    _lib.check(rc, 'flash_attn_bwd')
    return dqkv


def slot_stream_fwd(tokens, g, G, c0, want_attn=True, want_stats=True, eps=1e-5):
    """tokens fp32 or bf16 [B,N,768]; g [B,HS,768]; G,c0 [B,HS] -> U [B,HS,768], m, A [B,HS], attn [B,HS,N]|None, mu, rstd [B,N]|None
    (fp32 tokens: the fp32 streaming kernels of csrc/slot_attn.cu; bf16 tokens: the tcgen05 kernel of csrc/slot_attn_tc.cu)"""
    _need_cuda(tokens, g, G, c0)
    assert tokens.dtype in (torch.float32, torch.bfloat16), 'slot_stream_fwd: tokens must be fp32 or bf16'
    assert tokens.is_contiguous() and g.is_contiguous() and G.is_contiguous() and c0.is_contiguous()
    assert g.dtype == torch.float32 and G.dtype == torch.float32 and c0.dtype == torch.float32
    B, N, D = tokens.shape
    HS = g.shape[1]
    dev = tokens.device
    acc = torch.zeros(B * HS * (D + 2), device=dev, dtype=torch.float32)      # U | m | A accumulate (+=): one zero-fill
    U = acc[:B * HS * D].view(B, HS, D)
    mA = acc[B * HS * D:].view(2, B, HS)
    attn = torch.empty(B, HS, N, device=dev, dtype=torch.float32) if want_attn else None
    mu = torch.empty(B, N, device=dev, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(B, N, device=dev, dtype=torch.float32) if want_stats else None
    fn = _lib.lib().devias_slot_stream_fwd if tokens.dtype == torch.float32 else _lib.lib().devias_slot_stream_fwd_bf16
    rc = fn(tokens.data_ptr(), g.data_ptr(), G.data_ptr(), c0.data_ptr(), U.data_ptr(), mA[0].data_ptr(), mA[1].data_ptr(),
            _ptr(attn), _ptr(mu), _ptr(rstd), B, N, D, HS // 4, float(eps), _stream())
    _lib.check(rc, 'slot_stream_fwd')
    return U, mA[0], mA[1], attn, mu, rstd


SLOT_BF16_TOKENS = True     # bf16 context tokens: forward on the tcgen05 kernel (csrc/slot_attn_tc.cu)


def slot_stream_bwd(tokens, mu, rstd, g, G, attn, dU, dm, dA, dattn=None, dtokens=None):
    """-> (dtokens fp32, dg, dG, dc0); when `dtokens` is given the token gradient is accumulated into it in place.
    bf16 tokens: the tcgen05 backward of csrc/slot_attn_tc_bwd.cu (the token gradient stays fp32)."""
    _need_cuda(tokens, g, dU)
    assert tokens.dtype in (torch.float32, torch.bfloat16) and tokens.is_contiguous()
    B, N, D = tokens.shape
    HS = g.shape[1]
    dev = tokens.device
    acc = dtokens is not None
    if dtokens is None:
        dtokens = torch.empty(tokens.shape, device=dev, dtype=torch.float32)
    dg = torch.zeros(B, HS, D, device=dev, dtype=torch.float32)
    dGc = torch.zeros(2, B, HS, device=dev, dtype=torch.float32)
    cont = lambda t: None if t is None else t.contiguous()
    dU, dm, dA, dattn = cont(dU), cont(dm), cont(dA), cont(dattn)
    fn = _lib.lib().devias_slot_stream_bwd if tokens.dtype == torch.float32 else _lib.lib().devias_slot_stream_bwd_bf16
    rc = fn(tokens.data_ptr(), mu.data_ptr(), rstd.data_ptr(), g.data_ptr(), G.data_ptr(), attn.data_ptr(), dU.data_ptr(),
            dm.data_ptr(), dA.data_ptr(), _ptr(dattn), dtokens.data_ptr(), int(acc), dg.data_ptr(), dGc[0].data_ptr(),
            dGc[1].data_ptr(), B, N, D, HS // 4, _stream())
    _lib.check(rc, 'slot_stream_bwd')
    return dtokens, dg, dGc[0], dGc[1]
