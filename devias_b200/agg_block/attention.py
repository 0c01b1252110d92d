"""Parameter-owning modules of the aggregation block with the reference's interface (agg_block/attention.py: `cache_fn`,
`PreNorm`, `PostNorm`, `FeedForward`, `Attention`): same constructor arguments, same parameter names and shapes (state_dict
parity), same return values.

On the DEVIAS recipes `AggregationBlock.forward` never calls these modules' `forward`: it reads their parameters and runs the
folded streaming path (devias_b200/slot_attention.py + csrc/slot_attn.cu).  The `forward` methods below serve the remaining
configurations (post-norm, dropout, relu, direct use of a module) with plain tensor algebra written head-major
(`[batch, head, row, dim]`) instead of the reference's `(b h)` merges.
"""
import torch
import torch.nn.functional as F
from torch import nn


def exists(val):
    return val is not None


def default(val, d):
    return d if val is None else val


class _Memo:
    """Weight tying (agg_block/attention.py:12-23): a factory wrapped by `cache_fn` builds its module once and hands the same
    instance back on every later call, unless the call says `_cache=False`."""

    def __init__(self, factory):
        self.factory = factory
        self.instance = None
        self.__name__ = getattr(factory, '__name__', 'cached_fn')
        self.__doc__ = getattr(factory, '__doc__', None)

    def __call__(self, *args, _cache=True, **kwargs):
        if not _cache:
            return self.factory(*args, **kwargs)
        if self.instance is None:
            self.instance = self.factory(*args, **kwargs)
        return self.instance


def cache_fn(f):
    return _Memo(f)


class PreNorm(nn.Module):
    """LayerNorm on the input (and, when `context_dim` is given, on the `context` keyword) ahead of `fn`
    (agg_block/attention.py:25-40).  Attribute names `fn`, `norm`, `norm_context` are part of the state_dict."""

    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = None if context_dim is None else nn.LayerNorm(context_dim)

    def forward(self, x, **kwargs):
        if self.norm_context is not None:
            kwargs = dict(kwargs, context=self.norm_context(kwargs['context']))
        return self.fn(self.norm(x), **kwargs)


class PostNorm(nn.Module):
    """agg_block/attention.py:42-48"""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x)


_ACTIVATIONS = {'relu': nn.ReLU, 'gelu': nn.GELU}


class FeedForward(nn.Module):
    """Linear(dim -> mult*dim), activation, Dropout, Linear(mult*dim -> dim), Dropout | Identity as `net.0 .. net.4`
    (agg_block/attention.py:50-82; the signature's default 'geglu' is not implemented there either)."""

    def __init__(self, dim, mult=4, dropout=0., activation='geglu', more_dropout=False, xavier_init=False):
        super().__init__()
        if activation not in _ACTIVATIONS:
            raise NotImplementedError("Invalid activation function")
        hidden = int(dim * mult)
        self.activation = _ACTIVATIONS[activation]()
        tail = nn.Dropout(dropout) if more_dropout else nn.Identity()
        self.net = nn.Sequential(nn.Linear(dim, hidden), self.activation, nn.Dropout(dropout), nn.Linear(hidden, dim), tail)
        if xavier_init:
            self._reset_parameter()

    def _reset_parameter(self):
        for layer in self.net:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_normal_(layer.weight)
                nn.init.zeros_(layer.bias)

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    """Slot cross-attention (agg_block/attention.py:85-141).  The softmax runs over the QUERY (slot) axis -- the slots compete
    for every context token -- and the result is renormalised over the tokens (+1e-7) before it weights the values.
    Returns (to_out(weighted values) [B, n, query_dim], the slot-axis softmax [(B*heads), n, m])."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0., more_dropout=False, xavier_init=False):
        super().__init__()
        context_dim = default(context_dim, query_dim)
        width = heads * dim_head
        self.heads = heads
        self.query_sfmax_scale = dim_head ** -0.5
        self.key_softmax = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, width, bias=False)
        self.to_k = nn.Linear(context_dim, width, bias=False)
        self.to_v = nn.Linear(context_dim, width, bias=False)
        self.attn_holder = nn.Identity()
        self.attn_matrix_dropout = nn.Dropout(dropout) if more_dropout else nn.Identity()
        self.to_out = nn.Sequential(nn.Linear(width, query_dim), nn.Dropout(dropout))
        if xavier_init:
            self._reset_parameter()

    def _reset_parameter(self):
        for proj in (self.to_q, self.to_k, self.to_v):
            nn.init.xavier_uniform_(proj.weight)

    def _heads_first(self, t):
        """[B, rows, heads*d] -> [B, heads, rows, d]"""
        b, rows, _ = t.shape
        return t.view(b, rows, self.heads, -1).transpose(1, 2)

    def forward(self, x, context=None, k_pos=None, q_pos=None):
        context = x if context is None else context
        queries = self._heads_first(self.to_q(x if q_pos is None else x + q_pos))
        keys = self._heads_first(self.to_k(context if k_pos is None else context + k_pos))
        values = self._heads_first(self.to_v(context))
        logits = torch.matmul(queries, keys.transpose(-1, -2)) * self.query_sfmax_scale      # [B, h, n slots, m tokens]
        compete = F.softmax(logits, dim=-2)                                                  # over the slots
        b, h, n, m = compete.shape
        sim_distill = compete.reshape(b * h, n, m)
        weights = self.attn_holder(sim_distill).view(b, h, n, m)
        weights = weights / (weights.sum(dim=-1, keepdim=True) + 1e-7)
        weights = self.attn_matrix_dropout(weights)
        mixed = torch.matmul(weights, values).transpose(1, 2).reshape(b, n, -1)              # [B, n, heads*d]
        return self.to_out(mixed), sim_distill
