"""Drop-in for the reference's model/modeling_slot.py: the timm-registered DEVIAS student
`slot_vit_base_patch16_224` (VideoMAE-style ViT-B/16 over 1568 tube tokens + slot aggregation block
+ unified action/scene head + mask predictor) with identical constructor keywords, attributes,
state_dict keys/shapes and return tuples -- computed by hand-written sm_100a kernels.

Differences that are deliberate and documented in DESIGN.md:
  * the encoder computes in bf16 on tensor cores with an fp32 residual stream (reference: fp16 autocast/DeepSpeed);
  * `pos_embed` (not a parameter, not in the state_dict -- same as the reference) lives on the device instead of
    being copied host->device every forward (model/modeling_slot.py:355);
  * CUDA only: there is no CPU path (the CPU oracle is oracle/devias_oracle.py).
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .agg_block import AggregationBlock
from .functional import EncoderBlockFn, LayerNormFn, PatchEmbedFn

try:  # the reference registers its constructors with timm (model/modeling_slot.py:416); do the same when timm exists
    from timm.models.registry import register_model
except Exception:  # timm absent: plain function
    def register_model(fn):
        return fn


def _cfg(url='', **kwargs):
    return {'url': url, 'num_classes': 400, 'input_size': (3, 224, 224), 'pool_size': None, 'crop_pct': .9,
            'interpolation': 'bicubic', 'mean': (0.5, 0.5, 0.5), 'std': (0.5, 0.5, 0.5), **kwargs}


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def get_sinusoid_encoding_table(n_position, d_hid):
    """model/modeling_slot.py:181-191 -- table[p, j] = sin/cos(p / 10000^(2*(j//2)/d_hid)), float64 -> fp32 [1, P, D]."""
    j = np.arange(d_hid)
    ang = np.arange(n_position, dtype=np.float64)[:, None] / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.tensor(ang, dtype=torch.float, requires_grad=False).unsqueeze(0)


class MLPHead(nn.Module):
    """model/modeling_slot.py:23-33"""

    def __init__(self, in_dim, out_dim, hidden_dim):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, out_dim)
        self.act = nn.ReLU()

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class DropPath(nn.Module):
    """Per-sample stochastic depth (model/modeling_slot.py:36-47).  Inside a Block the Bernoulli(keep)/keep factor
    is handed to the residual GEMM epilogue as a per-sample row scale instead of being applied elementwise."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def row_scale(self, batch, device):
        if not self.training or not self.drop_prob:
            return None
        keep = 1.0 - self.drop_prob
        return torch.floor(keep + torch.rand(batch, device=device, dtype=torch.float32)) / keep

    def forward(self, x):
        s = self.row_scale(x.shape[0], x.device)
        return x if s is None else x * s.view(-1, *([1] * (x.dim() - 1))).to(x.dtype)

    def extra_repr(self):
        return 'p={}'.format(self.drop_prob)


class Mlp(nn.Module):
    """Parameter holder for model/modeling_slot.py:50-67 (fc1 -> exact GELU -> fc2 -> Dropout(drop))."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)


class Attention(nn.Module):
    """Parameter holder for model/modeling_slot.py:70-117 (qkv without bias + q_bias / v_bias, proj)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0., attn_head_dim=None):
        super().__init__()
        self.num_heads = num_heads
        head_dim = attn_head_dim if attn_head_dim is not None else dim // num_heads
        all_head_dim = head_dim * num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, all_head_dim * 3, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(all_head_dim)) if qkv_bias else None
        self.v_bias = nn.Parameter(torch.zeros(all_head_dim)) if qkv_bias else None
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(all_head_dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)


class Block(nn.Module):
    """model/modeling_slot.py:120-152; forward = ONE fused autograd Function over the sm_100a kernels."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0., drop_path=0.,
                 init_values=None, act_layer=nn.GELU, norm_layer=nn.LayerNorm, attn_head_dim=None):
        super().__init__()
        if drop or attn_drop:
            raise NotImplementedError('drop_rate / attn_drop_rate > 0 are never used by the DEVIAS recipes '
                                      '(run_slot_finetuning.py:61-66 defaults 0) and are not fused')
        if init_values and init_values > 0:
            raise NotImplementedError('layer-scale (init_values > 0) is not used by DEVIAS (modeling_slot.py:238)')
        if qk_scale is not None or attn_head_dim is not None or act_layer is not nn.GELU:
            raise NotImplementedError('non-default qk_scale / attn_head_dim / activation')
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        self.gamma_1, self.gamma_2 = None, None
        self._zero_bias = None
        self.last_scale = None

    def gemm_weights(self):
        return [self.attn.qkv.weight, self.attn.proj.weight, self.mlp.fc1.weight, self.mlp.fc2.weight]

    def forward(self, x, return_attn=False, w16=None, below_scale=None, scales=None):
        """below_scale: the MLP-branch drop-path factor of the block below (see EncoderBlockFn); `last_scale` is this block's.
        scales: (s1, s2) drop-path factors drawn by the caller for the whole encoder at once (else drawn here)."""
        if return_attn:
            raise NotImplementedError('return_attn=True materialises 12x1568x1568 attention maps; the fused flash path does '
                                      'not (and the reference slot model cannot run with it either, SURVEY.md R8)')
        a = self.attn
        if a.q_bias is None:
            if self._zero_bias is None or self._zero_bias.device != x.device:
                self._zero_bias = torch.zeros(a.qkv.weight.shape[0] // 3, device=x.device)
            qb = vb = self._zero_bias
        else:
            qb, vb = a.q_bias, a.v_bias
        s1 = s2 = None
        if scales is not None:
            s1, s2 = scales
        elif isinstance(self.drop_path, DropPath):
            s1 = self.drop_path.row_scale(x.shape[0], x.device)
            s2 = self.drop_path.row_scale(x.shape[0], x.device)
        if w16 is None:
            w16 = tuple(ops.cast_bf16(w.detach()) for w in self.gemm_weights())
        self.last_scale = s2
        return EncoderBlockFn.apply(x, self.norm1.weight, self.norm1.bias, a.qkv.weight, qb, vb, a.proj.weight, a.proj.bias,
                                    self.norm2.weight, self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias,
                                    self.mlp.fc2.weight, self.mlp.fc2.bias, w16, s1, s2, a.num_heads, self.norm1.eps, below_scale)


class PatchEmbed(nn.Module):
    """model/modeling_slot.py:155-177.  `proj` is kept as an nn.Conv3d so that parameter names/shapes match; the
    computation is patchify + tcgen05 GEMM (+ bias + sin-cos table in the epilogue)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, num_frames=16, tubelet_size=2):
        super().__init__()
        img_size, patch_size = _pair(img_size), _pair(patch_size)
        self.tubelet_size = int(tubelet_size)
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * (num_frames // self.tubelet_size)
        if patch_size != (16, 16) or self.tubelet_size != 2:
            raise NotImplementedError('the patchify kernel is specialised for 2x16x16 tubes (DEVIAS / VideoMAE ViT-B/16)')
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=(self.tubelet_size, *patch_size),
                              stride=(self.tubelet_size, *patch_size))

    def forward(self, x, pos_table=None, w16=None, **kwargs):
        B, C, T, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        if not x.is_cuda:
            raise RuntimeError('devias_b200 runs on CUDA only (no CPU fallback; the CPU oracle lives in oracle/)')
        D = self.proj.weight.shape[0]
        n = (T // 2) * (H // 16) * (W // 16)
        if pos_table is None:
            pos_table = torch.zeros(n, D, device=x.device)
        if w16 is None:
            w16 = ops.cast_bf16(self.proj.weight.detach())
        return PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, w16, pos_table)


class MaskPredictor(nn.Module):
    """model/modeling_slot.py:194-216"""

    def __init__(self):
        super().__init__()
        self.decoder = nn.Sequential(nn.Linear(768, 512), nn.ReLU(), nn.Linear(512, 256), nn.ReLU(),
                                     nn.Linear(256, 196), nn.Sigmoid())
        self.act = nn.ReLU()

    def forward(self, cls_token):
        if cls_token.is_cuda and cls_token.dtype == torch.float32:
            from . import slot_linear
            d = self.decoder
            h = F.relu(slot_linear.linear(cls_token, d[0].weight, d[0].bias))
            h = F.relu(slot_linear.linear(h, d[2].weight, d[2].bias))
            mask = torch.sigmoid(slot_linear.linear(h, d[4].weight, d[4].bias))
        else:
            mask = self.decoder(cls_token)
        return mask.squeeze().reshape(cls_token.shape[0], 14 * 14)


class _WeightArena:
    """All GEMM weights of the encoder as views of ONE flat fp32 buffer plus a flat bf16 shadow that the tensor-core
    kernels read; refreshing the shadow is a single cast launch (SURVEY.md section 8b 'private bf16 copies')."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8
        self.flat32 = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat16 = torch.empty(total, device=dev, dtype=torch.bfloat16)
        self.views16 = []
        with torch.no_grad():
            for p, o in zip(self.params, offs):
                v = self.flat32[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                self.views16.append(self.flat16[o:o + p.numel()].view(p.shape))
        self.ptrs = [p.data_ptr() for p in self.params]
        self.stamp = None

    def valid(self):
        return all(p.data_ptr() == q for p, q in zip(self.params, self.ptrs))

    def refresh(self, force):
        stamp = sum(p._version for p in self.params)
        if force or stamp != self.stamp:
            ops.cast_bf16(self.flat32, self.flat16)
            self.stamp = stamp


class VisionTransformer(nn.Module):
    """DEVIAS student (model/modeling_slot.py:219-413)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, fc_drop_rate=0., drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0., use_learnable_pos_emb=False, init_scale=0.,
                 all_frames=16, tubelet_size=2, use_checkpoint=False, num_latents=4, head_type='linear',
                 slot_matching_method='hard_select', num_scene_classes=365, agg_weights_tie=False, agg_depth=4,
                 slot_matching=None):
        super().__init__()
        if slot_matching is not None:  # run_slot_finetuning.py:386 passes this (misspelt) keyword; accepted, ignored
            pass
        if embed_dim != 768:
            raise NotImplementedError('kernels are instantiated for embed_dim = 768 (the slot head hard-codes 768 too, '
                                      'model/modeling_slot.py:392)')
        if use_checkpoint:
            raise NotImplementedError('use_checkpoint is broken in the reference (SURVEY.md R8) and not provided')
        self.num_slots = num_latents
        self.num_classes = num_classes
        self.num_scene_classes = num_scene_classes
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      num_frames=all_frames, tubelet_size=tubelet_size)
        num_patches = self.patch_embed.num_patches
        self.use_checkpoint = use_checkpoint
        self.slot_matching_method = slot_matching_method
        self.head_type = head_type
        self.select_slots_info = [[0, 0] for _ in range(self.num_slots)]
        if slot_matching_method not in ('hard_select', 'matching'):
            raise ValueError("incorrent slot_matching_method")

        if use_learnable_pos_emb:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        else:
            self.pos_embed = get_sinusoid_encoding_table(num_patches, embed_dim)   # plain attribute, as in the reference
        self.pos_drop = nn.Dropout(p=drop_rate)

        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, init_values=init_values)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        print(f"Aggregation blocks {agg_weights_tie} depth {agg_depth}")
        self.agg_block = AggregationBlock(num_latents=num_latents, weight_tie_layers=agg_weights_tie, depth=agg_depth)
        self.mask_predictor = MaskPredictor()
        if use_learnable_pos_emb:
            nn.init.trunc_normal_(self.pos_embed, std=.02)

        n_out = num_classes + self.num_scene_classes
        if head_type == 'linear':
            self.head = nn.Linear(embed_dim, n_out) if num_classes > 0 else nn.Identity()
            self.apply(self._init_weights)
            if num_classes > 0:
                self.head.weight.data.mul_(init_scale)
                self.head.bias.data.mul_(init_scale)
        else:
            self.head = MLPHead(embed_dim, n_out, hidden_dim=512) if num_classes > 0 else nn.Identity()
            self.apply(self._init_weights)
            if num_classes > 0:
                self.head.fc2.weight.data.mul_(init_scale)
                self.head.fc2.bias.data.mul_(init_scale)
        self._arena = None
        self._pos_dev = None
        #: dtype of the encoder tokens handed to the aggregation block (fp32 keeps the slot path at fp32 accuracy)
        # dtype of the encoder tokens handed to the aggregation block: fp32 (default; the fp32 streaming slot kernels, 1e-5
        # contract) or bf16 (what the reference's K/V projections see under autocast; the tcgen05 slot kernels).  Slots stay fp32.
        self.token_dtype = torch.float32

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def get_num_layers(self):
        return len(self.blocks)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = nn.Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def get_select_slot_info(self):
        print("action slot : " + " | ".join(str(i[0]) for i in self.select_slots_info))
        print("scene slot : " + " | ".join(str(i[1]) for i in self.select_slots_info))

    def reset_select_slot_info(self):
        self.select_slots_info = [[0, 0] for _ in range(self.num_slots)]

    # ------------------------------------------------------------------------------------------
    def invalidate_weight_cache(self):
        """call after modifying weights through `.data` (which does not bump tensor versions)"""
        if self._arena is not None:
            self._arena.stamp = None
        if self.__dict__.get('_param_arena') is not None:
            self._param_arena.invalidate16()

    def _weights16(self):
        gemm_params = [self.patch_embed.proj.weight] + [w for b in self.blocks for w in b.gemm_weights()]
        pa = self.__dict__.get('_param_arena')
        if pa is not None and pa.valid() and all(pa.contains(p) for p in gemm_params):
            # training arena (devias_b200/arena.py): the optimizer pass keeps the bf16 shadow fresh; re-cast only when someone
            # modified parameters through torch (load_state_dict, a torch optimizer)
            pa.refresh16()
            v = [pa.view16(p) for p in gemm_params]
            return v[0], [tuple(v[1 + 4 * i: 5 + 4 * i]) for i in range(len(self.blocks))]
        if self._arena is None or not self._arena.valid() or self._arena.params[0].device != gemm_params[0].device:
            self._arena = _WeightArena(gemm_params)
        self._arena.refresh(force=self.training and torch.is_grad_enabled())
        v = self._arena.views16
        return v[0], [tuple(v[1 + 4 * i: 5 + 4 * i]) for i in range(len(self.blocks))]

    def _draw_drop_path(self, batch, device):
        """Per-sample stochastic-depth factors floor(keep + U) / keep (model/modeling_slot.py:36-47) of ALL blocks and both
        branches in three launches instead of six per block; returns [(s1, s2)] per block (None where the rate is 0)."""
        if not self.training:
            return None
        rates = [float(getattr(b.drop_path, 'drop_prob', 0.0) or 0.0) for b in self.blocks]
        if not any(r > 0 for r in rates):
            return None
        key = (tuple(rates), str(device))
        if getattr(self, '_dp_keep', None) is None or self._dp_keep[0] != key:
            keep = torch.tensor([1.0 - r for r in rates for _ in range(2)], device=device, dtype=torch.float32).unsqueeze(1)
            self._dp_keep = (key, keep)
        keep = self._dp_keep[1]
        s = torch.floor(keep + torch.rand(keep.shape[0], batch, device=device, dtype=torch.float32)) / keep
        return [(s[2 * i], s[2 * i + 1]) if r > 0 else (None, None) for i, r in enumerate(rates)]

    def _pos_table(self, device):
        pe = self.pos_embed
        if isinstance(pe, nn.Parameter):
            return pe[0]
        if self._pos_dev is None or self._pos_dev.device != device:
            self._pos_dev = pe[0].to(device).contiguous()
        return self._pos_dev

    def forward_features(self, x, return_attn=False):
        if return_attn:
            raise NotImplementedError('return_attn=True is not provided (see Block.forward)')
        if not x.is_cuda:
            raise RuntimeError('devias_b200 runs on CUDA only (no CPU fallback; the CPU oracle lives in oracle/)')
        with torch.autocast('cuda', enabled=False):
            pe16, blk16 = self._weights16()
            pos = self._pos_table(x.device)
            if isinstance(self.pos_embed, nn.Parameter):
                x = self.patch_embed(x, pos_table=None, w16=pe16) + self.pos_embed
            else:
                x = self.patch_embed(x, pos_table=pos, w16=pe16)
            below = None          # drop-path factor of the consumer of each block's input gradient (fused into its LN backward)
            all_scales = self._draw_drop_path(x.shape[0], x.device)
            for i, (blk, w16) in enumerate(zip(self.blocks, blk16)):
                x = blk(x, w16=w16, below_scale=below, scales=None if all_scales is None else all_scales[i])
                below = blk.last_scale
            return LayerNormFn.apply(x, self.norm.weight, self.norm.bias, self.norm.eps, self.token_dtype, below)

    def _head_linear(self, x):
        if isinstance(self.head, nn.Linear) and x.dtype == torch.float32:
            from . import slot_linear
            return slot_linear.linear(x, self.head.weight, self.head.bias)
        return self.head(x)

    def forward(self, x, return_attn=False):
        x = self.forward_features(x, return_attn)
        self.agg_block.slot_dtype = torch.float32
        slots, attn = self.agg_block(x)
        with torch.autocast('cuda', enabled=False):
            if self.slot_matching_method == 'hard_select':
                action_feat, scene_feat = slots[:, 0], slots[:, 1]
                action_logit = self.head(self.fc_dropout(action_feat))
                scene_logit = self.head(self.fc_dropout(scene_feat))
                return (action_feat, scene_feat), (action_logit, scene_logit, []), ([], [], [])
            bs, num_slots, _ = slots.size()
            slots = slots.reshape(-1, 768)
            slots_head = self._head_linear(self.fc_dropout(slots))
            C = self.num_classes
            if slots_head.dtype == torch.float32 and num_slots <= 8 and slots_head.shape[-1] >= C + self.num_scene_classes:
                a_idx, s_idx = ops.slot_select(slots_head.detach(), num_slots, C, self.num_scene_classes)
            else:
                probs = F.softmax(slots_head, dim=-1).view(bs, num_slots, -1)
                a_idx = torch.argmax(probs[:, :, :C].max(dim=-1).values, dim=1)
                s_idx = torch.argmax(probs[:, :, C:C + self.num_scene_classes].max(dim=-1).values, dim=1)
            ar = torch.arange(bs, device=slots.device)
            s3, h3 = slots.view(bs, num_slots, -1), slots_head.view(bs, num_slots, -1)
            mask_predictions = self.mask_predictor(slots)
            return (s3[ar, a_idx], s3[ar, s_idx]), (h3[ar, a_idx], h3[ar, s_idx], attn), (slots_head, slots, mask_predictions)


@register_model
def slot_vit_base_patch16_224(pretrained=False, **kwargs):
    """model/modeling_slot.py:416-422"""
    model = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model
