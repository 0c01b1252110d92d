"""Module interface of the reference's agg_block/attention.py (PreNorm, PostNorm, FeedForward, Attention,
cache_fn) with identical constructor signatures, parameter names and return values.

These classes own the parameters (state_dict parity).  When they are driven through
`AggregationBlock.forward` on CUDA the work is done by the folded streaming slot-attention path
(devias_b200/slot_attention.py); calling a module directly keeps the reference semantics op by op.
"""
from functools import wraps

import torch
from torch import nn


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else d


def cache_fn(f):
    """agg_block/attention.py:12-23 -- weight tying: the first constructed module is returned again."""
    cache = None

    @wraps(f)
    def cached_fn(*args, _cache=True, **kwargs):
        if not _cache:
            return f(*args, **kwargs)
        nonlocal cache
        if cache is not None:
            return cache
        cache = f(*args, **kwargs)
        return cache
    return cached_fn


class PreNorm(nn.Module):
    """agg_block/attention.py:25-40"""

    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if exists(context_dim) else None

    def forward(self, x, **kwargs):
        x = self.norm(x)
        if exists(self.norm_context):
            kwargs.update(context=self.norm_context(kwargs['context']))
        return self.fn(x, **kwargs)


class PostNorm(nn.Module):
    """agg_block/attention.py:42-48"""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x)


class FeedForward(nn.Module):
    """agg_block/attention.py:50-82 (the default activation 'geglu' is rejected there too)"""

    def __init__(self, dim, mult=4, dropout=0., activation='geglu', more_dropout=False, xavier_init=False):
        super().__init__()
        act_in_dim = int(dim * mult)
        if activation == 'relu':
            self.activation = nn.ReLU()
        elif activation == 'gelu':
            self.activation = nn.GELU()
        else:
            raise NotImplementedError("Invalid activation function")
        self.net = nn.Sequential(
            nn.Linear(dim, act_in_dim),
            self.activation,
            nn.Dropout(dropout),
            nn.Linear(act_in_dim, dim),
            nn.Dropout(dropout) if more_dropout else nn.Identity(),
        )
        if xavier_init:
            self._reset_parameter()

    def _reset_parameter(self):
        def fn(m):
            if type(m) == nn.Linear:
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0.0)
        self.net.apply(fn)

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    """Slot cross-attention, agg_block/attention.py:85-141: softmax over the SLOT axis, token-axis
    renormalisation with +1e-7, returns (to_out(attn.v), slot-softmax before renormalisation)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0., more_dropout=False, xavier_init=False):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = default(context_dim, query_dim)
        self.query_sfmax_scale = dim_head ** -0.5
        self.key_softmax = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.attn_holder = nn.Identity()
        self.attn_matrix_dropout = nn.Dropout(dropout) if more_dropout else nn.Identity()
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        if xavier_init:
            self._reset_parameter()

    def _reset_parameter(self):
        nn.init.xavier_uniform_(self.to_q.weight)
        nn.init.xavier_uniform_(self.to_k.weight)
        nn.init.xavier_uniform_(self.to_v.weight)

    def forward(self, x, context=None, k_pos=None, q_pos=None):
        h = self.heads
        q = self.to_q(x if q_pos is None else x + q_pos)
        context = default(context, x)
        k = self.to_k(context if k_pos is None else context + k_pos)
        v = self.to_v(context)

        def split(t):  # 'b n (h d) -> (b h) n d'
            b, n, _ = t.shape
            return t.reshape(b, n, h, -1).permute(0, 2, 1, 3).reshape(b * h, n, -1)

        q, k, v = map(split, (q, k, v))
        sim = torch.einsum('bid,bjd->bij', q, k) * self.query_sfmax_scale
        attn = sim.softmax(dim=1)
        sim_distill = attn
        attn = self.attn_holder(attn)
        attn = attn / (attn.sum(dim=-1, keepdim=True) + 1e-7)
        attn = self.attn_matrix_dropout(attn)
        out = torch.einsum('bij,bjd->bid', attn, v)
        bh, n, d = out.shape
        out = out.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, h * d)
        return self.to_out(out), sim_distill
