"""Parameter-owning modules of the aggregation block with the reference's interface (agg_block/attention.py: `cache_fn`,
`PreNorm`, `PostNorm`, `FeedForward`, `Attention`): same constructor arguments, same parameter names and shapes (state_dict
parity), same return values.

On the DEVIAS recipes `AggregationBlock.forward` never calls these modules' `forward`: it reads their parameters and runs the
folded streaming path (devias_b200/slot_attention.py + csrc/slot_attn.cu).  Called directly with fp32 CUDA tensors the modules
run on the same hand-written kernels: `PreNorm(Attention)` = the folded streaming layer, a bare `Attention` = the folded
products against the un-normalised context (csrc/skinny.cu), `FeedForward` / `PreNorm(FeedForward)` = the slot-row products.
Only the configurations no DEVIAS recipe uses (dropout > 0 in training, non-fp32 inputs, host tensors in unit tests of the
module algebra) evaluate the plain tensor algebra below, written head-major (`[batch, head, row, dim]`).
"""
import torch
import torch.nn.functional as F
from torch import nn


def exists(val):
    return val is not None


def _kernel_ok(t):
    """fp32 CUDA tensor: the hand-written kernels apply"""
    return t is not None and torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32


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
        ctx = kwargs.get('context')
        if _kernel_ok(x):
            from .. import slot_kernels, slot_linear
            from .. import slot_attention as SA
            if (isinstance(self.fn, Attention) and self.norm_context is not None and _kernel_ok(ctx) and ctx.shape[-1] == 768
                    and kwargs.get('k_pos') is None and kwargs.get('q_pos') is None and self.fn._plain()):
                a = self.fn               # the whole PreNorm(Attention) layer folded onto ONE pass over the context tokens
                p = dict(norm=self.norm, norm_w=self.norm.weight, norm_b=self.norm.bias, ctx_w=self.norm_context.weight,
                         ctx_b=self.norm_context.bias, wq=a.to_q.weight, wk=a.to_k.weight, wv=a.to_v.weight,
                         wo=a.to_out[0].weight, bo=a.to_out[0].bias)
                with torch.autocast('cuda', enabled=False):
                    return SA.slot_attention_layer(x, ctx.reshape(ctx.shape[0], -1, ctx.shape[-1]), None, None, p,
                                                   stream=slot_kernels.slot_stream, lin=slot_linear)
            if self.norm_context is None and x.shape[-1] == 768:
                with torch.autocast('cuda', enabled=False):
                    return self.fn(slot_linear.layer_norm(x, self.norm), **kwargs)
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
        drop = self.net[2].p if isinstance(self.net[2], nn.Dropout) else 0.
        if _kernel_ok(x) and x.shape[-1] % 4 == 0 and (drop == 0. or not self.training) and not isinstance(self.net[4], nn.Dropout):
            from .. import slot_linear
            with torch.autocast('cuda', enabled=False):
                h = self.activation(slot_linear.linear(x, self.net[0].weight, self.net[0].bias))
                return slot_linear.linear(h, self.net[3].weight, self.net[3].bias)
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

    def _plain(self):
        """no dropout is active: the fused paths apply"""
        drop = self.to_out[1].p
        return (drop == 0. or not self.training) and not isinstance(self.attn_matrix_dropout, nn.Dropout)

    def _forward_kernels(self, x, context, k_pos, q_pos):
        """The cross-attention folded onto the slot rows (same algebra as devias_b200/slot_attention.py without the context
        LayerNorm): sim = (scale Wk_h^T q) . ctx_j and out = Wv_h (sum_j w_j ctx_j), every product on csrc/skinny.cu."""
        from .. import slot_linear as L
        b, n, _ = x.shape
        h = self.heads
        ctx = context.reshape(b, -1, context.shape[-1])
        dh = self.to_q.weight.shape[0] // h
        q = L.linear(x if q_pos is None else x + q_pos, self.to_q.weight).view(b, n, h, dh)
        qt = L.fold_keys(q, self.to_k.weight) * self.query_sfmax_scale                     # [B, h, n, D]
        keys_src = ctx if k_pos is None else ctx + k_pos
        logits = L.bmm_nt(qt.reshape(b, h * n, -1), keys_src).view(b, h, n, -1)             # [B, h, n, m]
        compete = F.softmax(logits, dim=-2)
        m = compete.shape[-1]
        sim_distill = compete.reshape(b * h, n, m)
        weights = compete / (compete.sum(dim=-1, keepdim=True) + 1e-7)
        cbar = L.bmm_nn(weights.reshape(b, h * n, m), ctx).view(b, h, n, -1)                # [B, h, n, D]
        mixed = L.apply_values(cbar, self.to_v.weight)                                      # [B, n, h*dh]
        return L.linear(mixed, self.to_out[0].weight, self.to_out[0].bias), sim_distill

    def forward(self, x, context=None, k_pos=None, q_pos=None):
        context = x if context is None else context
        if _kernel_ok(x) and _kernel_ok(context) and x.dim() == 3 and self._plain() and x.shape[-1] % 4 == 0 \
                and context.shape[-1] % 4 == 0:
            with torch.autocast('cuda', enabled=False):
                return self._forward_kernels(x, context, k_pos, q_pos)
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
