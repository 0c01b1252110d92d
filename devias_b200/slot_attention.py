"""Folded streaming formulation of DEVIAS slot attention (agg_block/attention.py:120-141 under
agg_block/attention.py:32-40 PreNorm), see DESIGN.md "slot attention".

Per head h and slot s the key/value projections are folded onto the (few) slot vectors:

    sim[h,s,j] = q~[h,s] . LN_ctx(t_j)                 q~ = scale * Wk_h^T (Wq_h LN(x_s))      in R^768
               = r_j * (g[h,s] . t_j - mu_j * G[h,s]) + c0[h,s]      g = q~ * gamma_c, G = sum(g), c0 = q~ . beta_c
    a = softmax over s;   A = sum_j a,  U = sum_j (a r_j) t_j,  m = sum_j a r_j mu_j
    out[s,h] = Wv_h ( (gamma_c * (U - m) + beta_c * A) / (A + 1e-7) )

so one layer streams the 1568x768 tokens ONCE (the `slot_stream` kernel) and the two 1568x768x2048
projections of the reference disappear.  Everything on the slot side is O(S) rows per clip.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

HEADS = 4


def token_stats(tokens: torch.Tensor, eps: float = 1e-5):
    """per-token LayerNorm statistics of the context (loop invariant over layers): mu, rstd [B, N]"""
    t = tokens.float()
    mu = t.mean(-1)
    var = (t - mu.unsqueeze(-1)).square().mean(-1)
    return mu, torch.rsqrt(var + eps)


def slot_stream_torch(tokens, mu, r, g, G, c0):
    """torch evaluation of the streaming step on the GPU (same contract as the CUDA kernels; used to differentiate the
    S = 8 micro-benchmark configuration, whose backward kernel is not instantiated, and by the kernel tests):
    tokens [B,N,D]; mu,r [B,N]; g [B,HS,D]; G,c0 [B,HS]  ->  U [B,HS,D], m [B,HS], A [B,HS], a [B,HS,N]"""
    t = tokens if tokens.dtype == torch.float64 else tokens.float()
    B, HS, D = g.shape
    S = HS // HEADS
    dots = torch.bmm(g, t.transpose(1, 2))                                   # [B,HS,N]
    sim = r.unsqueeze(1) * (dots - mu.unsqueeze(1) * G.unsqueeze(-1)) + c0.unsqueeze(-1)
    a = sim.view(B, HEADS, S, -1).softmax(dim=2).reshape(B, HS, -1)
    w = a * r.unsqueeze(1)
    U = torch.bmm(w, t)
    m = (w * mu.unsqueeze(1)).sum(-1)
    A = a.sum(-1)
    return U, m, A, a


def slot_attention_layer(x, tokens, mu, r, p, stream=slot_stream_torch, lin=None, **stream_kw):
    """One `PreNorm(Attention)` application on slots x [B,S,D] against tokens [B,N,D].
    p: dict with norm_w/b (slots LN), ctx_w/b (context LN), wq, wk, wv [2048,768], wo [768,2048], bo.
    lin: provider of the slot-row products (devias_b200.slot_linear on CUDA; torch expressions when None).
    Returns (to_out(attn.v) [B,S,D], sim_distill [(B*4), S, N])."""
    B, S, D = x.shape
    H = HEADS
    dh = p['wq'].shape[0] // H
    linear = F.linear if lin is None else lin.linear
    if lin is None:
        xn = F.layer_norm(x, (D,), p['norm_w'], p['norm_b'], 1e-5)
    else:
        xn = lin.layer_norm(x, p['norm'])
    q = linear(xn, p['wq']).view(B, S, H, dh)
    if lin is None:
        qt = torch.einsum('bshd,hdc->bhsc', q, p['wk'].view(H, dh, D)) * (dh ** -0.5)     # [B,H,S,D]
        g = (qt * p['ctx_w']).reshape(B, H * S, D)
        G = g.sum(-1)
        c0 = (qt @ p['ctx_b']).reshape(B, H * S)
    else:
        g, G, c0 = lin.fold_epilogue(lin.fold_keys(q, p['wk']), p['ctx_w'], p['ctx_b'], dh ** -0.5)
    U, m, A, a = stream(tokens, mu, r, g.contiguous(), G.contiguous(), c0.contiguous(), **stream_kw)
    if lin is None:
        cbar = (p['ctx_w'] * (U - m.unsqueeze(-1)) + p['ctx_b'] * A.unsqueeze(-1)) / (A.unsqueeze(-1) + 1e-7)
    else:
        cbar = lin.context(U, m, A, p['ctx_w'], p['ctx_b'], 1e-7)
    if lin is None:
        out = torch.einsum('bhsc,hdc->bshd', cbar.view(B, H, S, D), p['wv'].view(H, dh, D)).reshape(B, S, H * dh)
    else:
        out = lin.apply_values(cbar.view(B, H, S, D), p['wv'])
    return linear(out, p['wo'], p['bo']), a.reshape(B * H, S, -1)
