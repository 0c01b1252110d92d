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
