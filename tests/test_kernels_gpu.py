"""GPU parity of the memory-bound helper kernels against torch fp32 on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

from oracle import devias_oracle as O
from util import assert_close

pytestmark = pytest.mark.gpu


def _r(shape, seed, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(shape, device='cuda', generator=g) * scale


@pytest.mark.parametrize('rows', [1, 7, 1568, 3001])
def test_layernorm_fwd_bwd(rows):
    from devias_b200 import ops
    D = 768
    x = _r((rows, D), 1, 2.0) + 0.3
    g = 1 + 0.1 * _r((D,), 2); b = 0.1 * _r((D,), 3)
    for eps in (1e-6, 1e-5):
        ref = F.layer_norm(x, (D,), g, b, eps)
        y32, mean, rstd = ops.layernorm_fwd(x, g, b, eps, torch.float32)
        assert_close(y32, ref, 2e-6, 'ln fp32')
        y16, _, _ = ops.layernorm_fwd(x, g, b, eps, torch.bfloat16)
        assert_close(y16.float(), ref, 5e-3, 'ln bf16')
    # backward with residual add, bf16 copy and the three column reductions
    xr = x.clone().requires_grad_(True); gr = g.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
    dy = _r((rows, D), 4).bfloat16()
    dres = _r((rows, D), 5)
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy.float())
    dg = torch.zeros(D, device='cuda'); db = torch.zeros(D, device='cuda'); cs = torch.zeros(D, device='cuda')
    dx, dxb = ops.layernorm_bwd(dy, x, mean, rstd, g, d_resid=dres.clone(), dgamma=dg, dbeta=db, dx_colsum=cs)
    ref_dx = xr.grad + dres
    assert_close(dx, ref_dx, 5e-6, 'ln dx')
    assert_close(dxb.float(), ref_dx, 5e-3, 'ln dx bf16')
    assert_close(dg, gr.grad, 2e-5, 'ln dgamma')
    assert_close(db, br.grad, 2e-5, 'ln dbeta')
    assert_close(cs, ref_dx.sum(0), 2e-5, 'dx colsum')
    # fp32 dy, no residual, in-place semantics off
    dy32 = _r((rows, D), 6)
    xr.grad = None
    F.layer_norm(xr, (D,), gr, br, 1e-5).backward(dy32)
    dx2, _ = ops.layernorm_bwd(dy32, x, mean, rstd, g, want_bf16=False)
    assert_close(dx2, xr.grad, 5e-6, 'ln dx fp32 dy')


def test_colsum_and_casts():
    from devias_b200 import ops
    a = _r((3001, 2304), 1).bfloat16()
    out = torch.zeros(2304, device='cuda')
    ops.colsum_bf16(a, out)
    assert_close(out, a.float().sum(0), 1e-5, 'colsum')
    out2 = torch.zeros(768, device='cuda')
    ops.colsum_bf16(a[:, 1536:], out2)
    assert_close(out2, a[:, 1536:].float().sum(0), 1e-5, 'colsum strided view')
    x = _r((1000003,), 2)
    assert torch.equal(ops.cast_bf16(x), x.bfloat16())
    y = _r((4 * 50, 768), 3)
    s = torch.tensor([0., 1.25, 1.25, 0.], device='cuda')
    assert torch.equal(ops.scale_rows_cast(y, s, 50), (y * s.repeat_interleave(50)[:, None]).bfloat16())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float16])
def test_patchify_matches_conv3d_unfold(dtype):
    from devias_b200 import ops
    clip = O.synth_clips(2, seed=3).cuda().to(dtype)
    p = ops.patchify(clip)
    B, C, T, H, W = clip.shape
    ref = clip.float().reshape(B, C, T // 2, 2, H // 16, 16, W // 16, 16).permute(0, 2, 4, 6, 1, 3, 5, 7).reshape(B * 1568, 1536)
    assert torch.equal(p, ref.bfloat16())


def test_patch_embed_matches_oracle():
    from devias_b200 import ops
    sd = O.synth_state_dict(depth=0, seed=5)
    clip = O.synth_clips(2, seed=4)
    ref = O.patch_embed(sd, clip) + O.sinusoid_table(1568, 768)
    w16 = sd['patch_embed.proj.weight'].cuda().bfloat16().reshape(768, -1)
    x0 = ops.gemm(ops.patchify(clip.cuda()), w16, ops.EPI_RESID_F32, bias=sd['patch_embed.proj.bias'].cuda(),
                  aux=O.sinusoid_table(1568, 768)[0].cuda().contiguous(), aux_row_mod=1568)
    assert_close(x0.cpu().view(2, 1568, 768), ref, 4e-3, 'patch embed + pos')


@pytest.mark.parametrize('B', [1, 3])
def test_patch_embed_implicit_gemm_vs_conv3d(B):
    """the implicit-GEMM tube patch embedding (5-D TMA boxes over the NCTHW clip, tf32 MMAs) against the Conv3d of
    model/modeling_slot.py:167-177 + flatten/transpose + bias + position table, evaluated in float64"""
    from devias_b200 import ops
    clip = O.synth_clips(B, seed=21).cuda()
    w = _r((768, 3, 2, 16, 16), 6, 1.0 / 39.0)
    bias = _r((768,), 7, 0.02)
    pos = O.sinusoid_table(1568, 768)[0].cuda().contiguous()
    out = ops.patch_embed_fwd(clip, w, bias, pos)
    ref = F.conv3d(clip.double(), w.double(), bias.double(), stride=(2, 16, 16)).flatten(2).transpose(1, 2) + pos.double()
    assert out.shape == (B * 1568, 768)
    # tf32 operands (10-bit mantissa), fp32 accumulation over k = 1536
    assert_close(out.view(B, 1568, 768), ref, 1.5e-3, 'implicit patch embed')
