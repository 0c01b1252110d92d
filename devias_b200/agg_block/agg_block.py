"""AggregationBlock drop-in (reference: agg_block/agg_block.py:8-139): `num_latents` learned slot
vectors iteratively cross-attend to the encoder tokens; `depth` layers, optionally weight-tied.

Constructor keywords, parameter names and the (slots, sim) return value are the reference's.  On CUDA
the per-layer work runs through the folded streaming slot-attention path (one pass over the tokens per
layer); the nn.Modules created here hold the parameters and remain callable one by one.
"""
import torch
from torch import nn

from .. import slot_attention as SA
from .attention import Attention, FeedForward, PostNorm, PreNorm, cache_fn
from .pos_encoding import build_position_encoding


class AggregationBlock(nn.Module):
    def __init__(self, *, depth=4, input_channels=768, input_axis=2, num_latents=4, latent_dim=768, num_classes=1000,
                 attn_dropout=0., ff_dropout=0., weight_tie_layers=True, pos_enc_type='none', pre_norm=True,
                 post_norm=False, activation='gelu', last_ln=True, ff_mult=4, more_dropout=False, xavier_init=False,
                 query_fixed=False, query_xavier_init=False, query_type='learned', encoder_isab=False, first_order=False):
        super().__init__()
        self.input_axis = input_axis
        self.num_classes = num_classes
        self.input_dim = input_channels
        self.pos_enc = build_position_encoding(input_channels, pos_enc_type, input_axis)
        self.num_latents = num_latents
        self.query_type = query_type
        self.latent_dim = latent_dim
        self.encoder_isab = encoder_isab
        self.first_order = first_order
        self.depth = depth
        self.weight_tie_layers = weight_tie_layers
        self.slot_dtype = None      # None: slots come back in the context's dtype (reference behaviour); the student model pins fp32
        self._fast_ok = (pos_enc_type == 'none' and pre_norm and not post_norm and activation == 'gelu'
                         and attn_dropout == 0. and ff_dropout == 0. and not more_dropout)

        if query_type == 'learned':
            self.latents = nn.Parameter(torch.randn(num_latents, latent_dim))
            if query_fixed:
                self.latents.requires_grad = False
            if query_xavier_init:
                nn.init.xavier_normal_(self.latents)
        elif query_type == 'slot':
            gain = nn.init.calculate_gain('linear')
            self.slots_mu = nn.init.xavier_uniform_(nn.Parameter(torch.randn(1, 1, latent_dim)), gain=gain)
            self.slots_log_sigma = nn.init.xavier_uniform_(nn.Parameter(torch.randn(1, 1, latent_dim)), gain=gain)
        else:
            raise NotImplementedError

        assert (pre_norm or post_norm)
        wrap = PreNorm if pre_norm else (lambda dim, fn, context_dim=None: fn)
        post = PostNorm if post_norm else (lambda dim: nn.Identity())

        make_attn = cache_fn(lambda: wrap(
            latent_dim,
            Attention(latent_dim, input_channels, heads=4, dim_head=512, dropout=attn_dropout,
                      more_dropout=more_dropout, xavier_init=xavier_init),
            context_dim=input_channels))
        make_ff = cache_fn(lambda: wrap(
            latent_dim,
            FeedForward(latent_dim, dropout=ff_dropout, activation=activation, mult=ff_mult,
                        more_dropout=more_dropout, xavier_init=xavier_init)))

        self.layers = nn.ModuleList([])
        print(f"slot attention block : weight tie : {weight_tie_layers}, depth : {depth}")
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                make_attn(_cache=weight_tie_layers), post(latent_dim),
                make_ff(_cache=weight_tie_layers), post(latent_dim)]))
        self.last_layer = nn.Sequential(nn.LayerNorm(latent_dim) if last_ln and not post_norm else nn.Identity())
        self.encoder_output_holder = nn.Identity()
        self.decoder_output_holder = nn.Identity()

    def get_queries(self, b):
        if self.query_type == 'learned':
            return self.latents.unsqueeze(0).expand(b, -1, -1)
        init = torch.randn((b, self.num_latents, self.latent_dim), device=self.slots_mu.device)
        return self.slots_mu + self.slots_log_sigma.exp() * init

    # ------------------------------------------------------------------------------------------
    def _layer_params(self, cross_attn):
        a = cross_attn.fn
        return dict(norm=cross_attn.norm, norm_w=cross_attn.norm.weight, norm_b=cross_attn.norm.bias,
                    ctx_w=cross_attn.norm_context.weight, ctx_b=cross_attn.norm_context.bias,
                    wq=a.to_q.weight, wk=a.to_k.weight, wv=a.to_v.weight,
                    wo=a.to_out[0].weight, bo=a.to_out[0].bias)

    def _forward_streaming(self, data):
        from .. import slot_kernels, slot_linear
        with torch.autocast('cuda', enabled=False):
            x = self.get_queries(data.shape[0]).float()
            mu, r = slot_kernels.token_stats(data)
            sink = slot_kernels.new_sink() if (torch.is_grad_enabled() and data.requires_grad) else None
            sim = None
            for cross_attn, _, cross_ff, _ in self.layers:
                attn, sim = SA.slot_attention_layer(x, data, mu, r, self._layer_params(cross_attn),
                                                    stream=slot_kernels.slot_stream, lin=slot_linear, sink=sink)
                x = attn + x
                net = cross_ff.fn.net                            # PreNorm(FeedForward): Linear, GELU, Dropout(0), Linear
                h = slot_linear.linear(slot_linear.layer_norm(x, cross_ff.norm), net[0].weight, net[0].bias)
                x = slot_linear.linear(net[1](h), net[3].weight, net[3].bias) + x
            last = self.last_layer[0]
            out = slot_linear.layer_norm(x, last) if isinstance(last, nn.LayerNorm) else last(x)
            # the reference carries the slots in the context's dtype (agg_block/agg_block.py:128 `.type_as(data)`)
            want = data.dtype if self.slot_dtype is None else self.slot_dtype
            return (out if out.dtype == want else out.to(want)), sim

    def forward(self, data):
        b, *axis = data.shape    # as in the reference (agg_block/agg_block.py:121-122) the channel dim counts as an axis
        assert len(axis) == self.input_axis, 'input data must have the right number of axis'
        pos = self.pos_enc(data)
        data = data.reshape(b, -1, data.shape[-1])
        if not data.is_cuda:
            raise RuntimeError('devias_b200.AggregationBlock runs on CUDA only (no CPU fallback; the CPU oracle lives in oracle/)')
        if self._fast_ok and self.query_type == 'learned':
            return self._forward_streaming(data)
        # configurations no DEVIAS recipe uses (post-norm, dropout, relu, 'slot' queries): reference op order via torch
        x = self.get_queries(b).type_as(data)
        sim = None
        for cross_attn, pn1, cross_ff, pn2 in self.layers:
            attn, sim = cross_attn(x, context=data, k_pos=pos, q_pos=None)
            x = pn1(attn + x)
            x = pn2(cross_ff(x) + x)
        x = self.decoder_output_holder(x)
        return self.last_layer(x), sim
