"""Drop-in for the reference's model/modeling_finetune.py:178-334 -- the plain VideoMAE ViT-B/16 classifier that DEVIAS
uses as its FROZEN SCENE TEACHER (`vit_base_patch16_224(num_classes=365, use_mean_pooling=False)`: CLS token, 1569 tokens,
run under no_grad in every training step, engine/engine_for_slot.py:52-53).  SURVEY.md section 8f row N1.

Same encoder kernels as the student (devias_b200/modeling_slot.py); N = 1569 is handled by every kernel (ragged tiles).
"""
from functools import partial

import torch
import torch.nn as nn

from .functional import LayerNormFn
from .modeling_slot import Block, PatchEmbed, _WeightArena, _cfg, get_sinusoid_encoding_table, register_model


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, fc_drop_rate=0., drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0., use_learnable_pos_emb=False, init_scale=0.,
                 all_frames=16, tubelet_size=2, use_checkpoint=False, use_mean_pooling=True):
        super().__init__()
        if embed_dim != 768:
            raise NotImplementedError('kernels are instantiated for embed_dim = 768')
        if use_checkpoint:
            raise NotImplementedError('use_checkpoint is not provided')
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      num_frames=all_frames, tubelet_size=tubelet_size)
        num_patches = self.patch_embed.num_patches
        self.use_checkpoint = use_checkpoint
        if not use_mean_pooling:
            print("set cls")
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
            nn.init.trunc_normal_(self.cls_token, std=.02)
            num_patches += 1
        if use_learnable_pos_emb:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        else:
            self.pos_embed = get_sinusoid_encoding_table(num_patches, embed_dim)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, init_values=init_values)
            for i in range(depth)])
        self.norm = nn.Identity() if use_mean_pooling else norm_layer(embed_dim)
        self.fc_norm = norm_layer(embed_dim) if use_mean_pooling else None
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        self.head = nn.Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        if use_learnable_pos_emb:
            nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)
        if num_classes > 0:
            self.head.weight.data.mul_(init_scale)
            self.head.bias.data.mul_(init_scale)
        self._arena = None
        self._pos_dev = None
        self._zero_pos = None

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

    def _weights16(self):
        params = [self.patch_embed.proj.weight] + [w for b in self.blocks for w in b.gemm_weights()]
        if self._arena is None or not self._arena.valid() or self._arena.params[0].device != params[0].device:
            self._arena = _WeightArena(params)
        self._arena.refresh(force=self.training and torch.is_grad_enabled())
        v = self._arena.views16
        return v[0], [tuple(v[1 + 4 * i: 5 + 4 * i]) for i in range(len(self.blocks))]

    def forward_features(self, x, return_attn=False):
        if return_attn:
            raise NotImplementedError('return_attn=True is not provided (the fused attention does not materialise the maps)')
        if not x.is_cuda:
            raise RuntimeError('devias_b200 runs on CUDA only (no CPU fallback; the CPU oracle lives in oracle/)')
        with torch.autocast('cuda', enabled=False):
            pe16, blk16 = self._weights16()
            dev = x.device
            if self._pos_dev is None or self._pos_dev.device != dev:
                pe = self.pos_embed
                self._pos_dev = pe if isinstance(pe, nn.Parameter) else pe.to(dev)
            n_patch = self.patch_embed.num_patches
            if self.fc_norm is None:
                # CLS model: the position table has 1569 rows (row 0 = CLS), so it cannot ride in the patch-embed epilogue
                if self._zero_pos is None or self._zero_pos.device != dev:
                    self._zero_pos = torch.zeros(n_patch, self.embed_dim, device=dev)
                x = self.patch_embed(x, pos_table=self._zero_pos, w16=pe16)
                x = torch.cat((self.cls_token.expand(x.size(0), -1, -1), x), dim=1) + self._pos_dev
            elif isinstance(self.pos_embed, nn.Parameter):
                # learnable table (model/modeling_finetune.py:221-222): added differentiably so that it receives its gradient
                if self._zero_pos is None or self._zero_pos.device != dev:
                    self._zero_pos = torch.zeros(n_patch, self.embed_dim, device=dev)
                x = self.patch_embed(x, pos_table=self._zero_pos, w16=pe16) + self.pos_embed
            else:
                x = self.patch_embed(x, pos_table=self._pos_dev[0].contiguous(), w16=pe16)
            x = x.contiguous()
            for blk, w16 in zip(self.blocks, blk16):
                x = blk(x, w16=w16)
            if self.fc_norm is not None:
                return self.fc_norm(x.mean(1))
            cls = x[:, 0].contiguous()
            return LayerNormFn.apply(cls, self.norm.weight, self.norm.bias, self.norm.eps, torch.float32)

    def forward(self, x, return_attn=False):
        token = self.forward_features(x, return_attn)
        with torch.autocast('cuda', enabled=False):
            x = self.head(self.fc_dropout(token))
        return token, x


@register_model
def vit_base_patch16_224(pretrained=False, **kwargs):
    """model/modeling_finetune.py:328-334"""
    model = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model
