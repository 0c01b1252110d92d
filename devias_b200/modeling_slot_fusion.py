"""Drop-in for the reference's model/modeling_slot_fusion.py (downstream fine-tuning on the disentangled slots): same
constructor signature, parameter names / shapes (state_dict parity) and return values, on the sm_100a encoder, slot
aggregation, slot-selection and slot-row kernels of this package.

    forward(x) (model/modeling_slot_fusion.py:364-403)
      'gap'   : fusion_head(fc_dropout(action_norm(tokens.mean(1))))           -> (x.mean(1), x)   [reference quirk kept]
      'concat': slots = agg_block(tokens); the pre-trained `head` picks the action / scene slot (softmax, argmax over the
                slots); (action_norm(action slot), scene_norm(scene slot)) -> fusion_head -> (concat features, logits)

Reference quirks kept on purpose: `MLPHead.forward` sends BOTH tokens through fc_action_down / fc_action_ln (:43-44, the
scene branch's parameters exist but are unused), and with head_type='linear' + 'concat' the reference builds an nn.Linear but
calls it with two arguments (:398) -- a TypeError there and here.
"""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, slot_linear
from .agg_block import AggregationBlock
from .modeling_slot import Block, PatchEmbed, VisionTransformer as _SlotViT, _cfg, get_sinusoid_encoding_table, register_model


def _lin(x, layer):
    if x.is_cuda and x.dtype == torch.float32:
        return slot_linear.linear(x, layer.weight, layer.bias)
    return layer(x)


class MLPHead(nn.Module):
    """model/modeling_slot_fusion.py:23-54"""

    def __init__(self, in_dim, out_dim, fc_drop_rate=0., use_input_ln=True):
        super().__init__()
        self.fc_action_down = nn.Linear(in_dim, in_dim // 2)
        self.fc_scene_down = nn.Linear(in_dim, in_dim // 2)
        self.fc_action_ln = nn.LayerNorm(in_dim // 2)
        self.fc_scene_ln = nn.LayerNorm(in_dim // 2)
        self.use_input_ln = use_input_ln
        if use_input_ln:
            self.fc_input_ln = nn.LayerNorm(in_dim)
        self.classifier = nn.Linear(in_dim, out_dim)
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        self.relu = nn.ReLU()

    def forward(self, action_token, scene_token):
        action_token = self.fc_action_ln(_lin(action_token, self.fc_action_down))
        scene_token = self.fc_action_ln(_lin(scene_token, self.fc_action_down))     # sic: the action branch, as in the reference
        output = torch.concat([action_token, scene_token], dim=1)
        if self.use_input_ln:
            output = self.fc_input_ln(output)
        return _lin(self.fc_dropout(self.relu(output)), self.classifier)


class VisionTransformer(_SlotViT):
    """model/modeling_slot_fusion.py:186-403.  Encoder / weight-arena / positional-table plumbing is inherited from the
    pre-training model (devias_b200.modeling_slot.VisionTransformer); the module tree is the fusion model's."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=False, qk_scale=None, fc_drop_rate=0., drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=nn.LayerNorm, init_values=0., use_learnable_pos_emb=False, init_scale=0.,
                 all_frames=16, tubelet_size=2, use_checkpoint=False, num_latents=4, head_type='linear', agg_weights_tie=True,
                 agg_depth=4, num_scene_classes=365, slot_fusion_method='concat', downstream_nb_classes=50, use_input_ln=True):
        nn.Module.__init__(self)
        if embed_dim != 768:
            raise NotImplementedError('kernels are instantiated for embed_dim = 768')
        if use_checkpoint:
            raise NotImplementedError('use_checkpoint is not provided (SURVEY.md R8)')
        self.num_slots = num_latents
        self.num_classes = num_classes
        self.num_scene_classes = num_scene_classes
        self.num_features = self.embed_dim = embed_dim
        self.tubelet_size = tubelet_size
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      num_frames=all_frames, tubelet_size=tubelet_size)
        num_patches = self.patch_embed.num_patches
        self.use_checkpoint = use_checkpoint
        self.slot_fusion_method = slot_fusion_method
        if use_learnable_pos_emb:
            self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, embed_dim))
        else:
            self.pos_embed = get_sinusoid_encoding_table(num_patches, embed_dim)
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.use_input_ln = use_input_ln
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                  drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer, init_values=init_values)
            for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.fc_dropout = nn.Dropout(p=fc_drop_rate) if fc_drop_rate > 0 else nn.Identity()
        print(f"Aggregation blocks {agg_weights_tie} depth {agg_depth}")
        self.agg_block = AggregationBlock(num_latents=num_latents, weight_tie_layers=agg_weights_tie, depth=agg_depth)
        if use_learnable_pos_emb:
            nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.action_norm = norm_layer(embed_dim)
        self.scene_norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes + self.num_scene_classes)
        if head_type == 'linear':
            if slot_fusion_method == 'concat':
                self.fusion_head = nn.Linear(embed_dim * num_latents, downstream_nb_classes) if downstream_nb_classes > 0 else nn.Identity()
            elif slot_fusion_method == 'gap':
                self.fusion_head = nn.Linear(embed_dim, downstream_nb_classes) if downstream_nb_classes > 0 else nn.Identity()
            nn.init.trunc_normal_(self.fusion_head.weight, std=.02)
            self.apply(self._init_weights)
            self.fusion_head.weight.data.mul_(init_scale)
            self.fusion_head.bias.data.mul_(init_scale)
        else:
            if slot_fusion_method == 'concat':
                self.fusion_head = (MLPHead(embed_dim, downstream_nb_classes, fc_drop_rate=fc_drop_rate, use_input_ln=use_input_ln)
                                    if downstream_nb_classes > 0 else nn.Identity())
            else:
                raise NotImplementedError()
            nn.init.trunc_normal_(self.fusion_head.classifier.weight, std=.02)
            self.apply(self._init_weights)
        self._arena = None
        self._pos_dev = None
        self.token_dtype = torch.float32

    def forward(self, x, return_attn=False):
        x = self.forward_features(x, return_attn)
        with torch.autocast('cuda', enabled=False):
            if self.slot_fusion_method == 'gap':
                x = self.fc_dropout(self.action_norm(x.mean(1)))
                x = _lin(x, self.fusion_head) if isinstance(self.fusion_head, nn.Linear) else self.fusion_head(x)
                return x.mean(1), x
            slots, attn = self.agg_block(x)
            bs, num_slots, _ = slots.size()
            slots = slots.reshape(-1, 768)
            slots_head = _lin(slots, self.head)
            C, Cs = self.num_classes, self.num_scene_classes
            if slots_head.dtype == torch.float32 and num_slots <= 8:
                a_idx, s_idx = ops.slot_select(slots_head.detach(), num_slots, C, Cs)
            else:
                probs = F.softmax(slots_head, dim=-1).view(bs, num_slots, -1)
                a_idx = torch.argmax(probs[:, :, :C].max(dim=-1).values, dim=1)
                s_idx = torch.argmax(probs[:, :, C:C + Cs].max(dim=-1).values, dim=1)
            ar = torch.arange(bs, device=slots.device)
            s3 = slots.view(bs, num_slots, -1)
            action_feat = self.action_norm(s3[ar, a_idx])
            scene_feat = self.scene_norm(s3[ar, s_idx])
            if self.slot_fusion_method == 'concat':
                inp = torch.concat((action_feat, scene_feat), dim=1)
                output = self.fusion_head(action_feat, scene_feat)
                return inp, output
            raise ValueError('fusion error')


@register_model
def slot_fusion_vit_base_patch16_224(pretrained=False, **kwargs):
    """model/modeling_slot_fusion.py:406-411"""
    model = VisionTransformer(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(nn.LayerNorm, eps=1e-6), **kwargs)
    model.default_cfg = _cfg()
    return model
