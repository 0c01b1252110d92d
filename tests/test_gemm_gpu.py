"""GPU parity: tcgen05 GEMM (all layouts / epilogues) against torch fp32 matmul on the same bf16 inputs."""
import pytest
import torch

from util import assert_close

pytestmark = pytest.mark.gpu


def _ops():
    from devias_b200 import ops
    return ops


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return (torch.randn(shape, device='cuda', generator=g) * scale)


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (128, 128, 128), (1568, 768, 768), (3136, 2304, 768), (200, 3072, 768),
                                   (1568 * 2, 768, 3072), (96, 32, 8)])
def test_gemm_nt_store(M, N, K):
    ops = _ops()
    a = _rand((M, K), seed=1).bfloat16(); b = _rand((N, K), 0.05, seed=2).bfloat16(); bias = _rand((N,), seed=3)
    ref = a.float() @ b.float().t() + bias
    out = ops.gemm(a, b, ops.EPI_STORE_F32, bias=bias)
    assert_close(out, ref, 2e-5, 'f32 store')
    outb = ops.gemm(a, b, ops.EPI_STORE_BF16, bias=bias)
    assert_close(outb.float(), ref, 6e-3, 'bf16 store')
    out_nb = ops.gemm(a, b, ops.EPI_STORE_F32)
    assert_close(out_nb, ref - bias, 2e-5, 'no bias')


def test_gemm_gelu_and_dgelu():
    ops = _ops()
    M, N, K = 1568, 3072, 768
    a = _rand((M, K), seed=1).bfloat16(); b = _rand((N, K), 0.05, seed=2).bfloat16(); bias = _rand((N,), seed=3)
    pre = a.float() @ b.float().t() + bias
    o1, o2 = ops.gemm(a, b, ops.EPI_GELU_BF16, bias=bias)
    assert_close(o1.float(), pre, 6e-3, 'pre-activation')
    assert_close(o2.float(), torch.nn.functional.gelu(pre), 6e-3, 'gelu')
    # dgrad with GELU' epilogue: dh = (dy @ W2) * gelu'(h_pre);  W2 is [Nout, Nin] = [k, n] -> b_mn
    dy = _rand((M, 768), seed=5).bfloat16(); w2 = _rand((768, N), 0.05, seed=6).bfloat16()
    h = o1
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).backward(dy.float() @ w2.float())
    dh = ops.gemm(dy, w2, ops.EPI_DGELU_BF16, b_mn=True, aux=h)
    assert_close(dh.float(), hf.grad, 8e-3, 'dgelu')


def test_gemm_resid_rowscale_and_mod():
    ops = _ops()
    M, N, K = 1568 * 2, 768, 768
    a = _rand((M, K), seed=1).bfloat16(); b = _rand((N, K), 0.05, seed=2).bfloat16(); bias = _rand((N,), seed=3)
    resid = _rand((M, N), seed=4)
    scale = torch.tensor([0.0, 1.25], device='cuda')
    ref = resid + scale.repeat_interleave(1568)[:, None] * (a.float() @ b.float().t() + bias)
    out = ops.gemm(a, b, ops.EPI_RESID_F32, bias=bias, aux=resid, row_scale=scale, rows_per_scale=1568)
    assert_close(out, ref, 2e-5, 'resid+rowscale')
    pos = _rand((1568, N), seed=7)
    ref2 = pos.repeat(2, 1) + (a.float() @ b.float().t() + bias)
    out2 = ops.gemm(a, b, ops.EPI_RESID_F32, bias=bias, aux=pos, aux_row_mod=1568)
    assert_close(out2, ref2, 2e-5, 'resid row-mod (pos-embed)')
    # in-place residual update (out aliases aux), as the encoder uses it
    x = resid.clone()
    ops.gemm(a, b, ops.EPI_RESID_F32, bias=bias, aux=x, out=x)
    assert_close(x, resid + (a.float() @ b.float().t() + bias), 2e-5, 'in-place residual')


@pytest.mark.parametrize('M,N,K', [(1568, 768, 2304), (1568 * 3, 3072, 768), (1000, 768, 768)])
def test_gemm_dgrad_layout(M, N, K):
    """dX[M,N] = dY[M,K] @ W[K,N]: B given as [k][n] (b_mn)."""
    ops = _ops()
    dy = _rand((M, K), seed=1).bfloat16(); w = _rand((K, N), 0.05, seed=2).bfloat16()
    out = ops.gemm(dy, w, ops.EPI_STORE_F32, b_mn=True)
    assert_close(out, dy.float() @ w.float(), 2e-5, 'dgrad')


@pytest.mark.parametrize('T,Nout,Nin,split', [(1568, 768, 768, 1), (1568 * 2, 3072, 768, 4), (1568 * 3, 768, 3072, 8),
                                              (1568, 2304, 768, 3), (1000, 768, 768, 16)])
def test_gemm_wgrad_layout(T, Nout, Nin, split):
    """dW[Nout,Nin] += dY[T,Nout]^T @ X[T,Nin]: both operands [k][m|n] (a_mn, b_mn), split-K atomics."""
    ops = _ops()
    dy = _rand((T, Nout), seed=1).bfloat16(); x = _rand((T, Nin), seed=2).bfloat16()
    ref = dy.float().t() @ x.float()
    out = torch.zeros(Nout, Nin, device='cuda')
    ops.gemm(dy, x, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=out, split_k=split)
    assert_close(out, ref, 3e-5, 'wgrad')
    ops.gemm(dy, x, ops.EPI_ATOMIC_F32, a_mn=True, b_mn=True, out=out, split_k=split)
    assert_close(out, 2 * ref, 3e-5, 'wgrad accumulate')


def test_gemm_rejects_bad_args():
    ops = _ops()
    a = torch.zeros(128, 64, device='cuda', dtype=torch.bfloat16); b = torch.zeros(40, 64, device='cuda', dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.gemm(a, b, ops.EPI_STORE_F32)  # n % 32 != 0
    with pytest.raises(RuntimeError):
        ops.gemm(a.cpu(), b.cpu(), ops.EPI_STORE_F32)

