"""CPU oracle for the DEVIAS hot path (TEST INFRASTRUCTURE -- never shipped, never measured).

A plain torch-fp32, CPU-only *restatement* of the reference algorithm for the path
`slot_vit_base_patch16_224` -> encoder -> AggregationBlock -> head/slot matching.
It is written functionally over a ``state_dict`` (same keys/shapes as the reference
model) so the very same weights can be fed to the reference (in the build container),
to this oracle (everywhere) and to the CUDA product path (GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this module. The product package ``devias_b200`` must never import it.

Parity pinning: the reference ships no tests / golden vectors.  The oracle is
pinned against outputs of the *reference itself* executed in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``) and re-checked live against
``/root/reference`` whenever that tree is present (``tests/test_oracle_golden.py::test_oracle_vs_live_reference``).

Every function cites the reference file:line it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
State = Dict[str, Tensor]


# ----------------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs (shared by golden generation, tests, bench, smoke)
# ----------------------------------------------------------------------------------------------

def _trunc_normal(rs: np.random.RandomState, shape, std=0.02) -> Tensor:
    # timm/torch trunc_normal_(std=.02, a=-2, b=2): the +-2 cut is 100 sigma away -> plain normal
    # is distribution-identical to 1e-2000; we only need *a* deterministic weight set.
    return torch.from_numpy((rs.standard_normal(size=shape) * std).astype(np.float32))


def synth_state_dict(num_classes=101, num_scene_classes=365, num_latents=2, agg_depth=4,
                     agg_weights_tie=True, depth=12, embed_dim=768, mlp_ratio=4, seed=0,
                     bias_std=0.02, ln_jitter=0.1, head_std=0.02, tubelet=2, patch=16, in_chans=3) -> State:
    """Deterministic (numpy RandomState => platform-stable) state_dict with the reference's
    key set and shapes (model/modeling_slot.py:222-316, agg_block/agg_block.py:78-107).

    Unlike the reference init (zero biases, unit LayerNorm) every tensor is made non-trivial
    (biases ~ N(0, bias_std), LN weight 1 + N(0, ln_jitter), LN bias ~ N(0, bias_std)) so parity
    tests exercise every term.  Tied aggregation layers repeat the same tensors under every
    ``agg_block.layers.<l>`` key exactly as the reference state_dict does (attention.py:12-23).
    """
    rs = np.random.RandomState(seed)
    D, Hd = embed_dim, int(embed_dim * mlp_ratio)
    sd: State = {}

    def lin(prefix, out_f, in_f, bias=True, std=0.02):
        sd[prefix + '.weight'] = _trunc_normal(rs, (out_f, in_f), std)
        if bias:
            sd[prefix + '.bias'] = _trunc_normal(rs, (out_f,), bias_std)

    def ln(prefix, dim):
        sd[prefix + '.weight'] = 1.0 + _trunc_normal(rs, (dim,), ln_jitter)
        sd[prefix + '.bias'] = _trunc_normal(rs, (dim,), bias_std)

    K = in_chans * tubelet * patch * patch
    sd['patch_embed.proj.weight'] = _trunc_normal(rs, (D, in_chans, tubelet, patch, patch), 1.0 / math.sqrt(K))
    sd['patch_embed.proj.bias'] = _trunc_normal(rs, (D,), bias_std)
    for i in range(depth):
        p = f'blocks.{i}.'
        ln(p + 'norm1', D)
        sd[p + 'attn.q_bias'] = _trunc_normal(rs, (D,), bias_std)
        sd[p + 'attn.v_bias'] = _trunc_normal(rs, (D,), bias_std)
        # std chosen so activations stay O(1) through 12 blocks (like a trained net, unlike 0.02 init)
        sd[p + 'attn.qkv.weight'] = _trunc_normal(rs, (3 * D, D), 0.04)
        lin(p + 'attn.proj', D, D, std=0.02)
        ln(p + 'norm2', D)
        lin(p + 'mlp.fc1', Hd, D, std=0.03)
        lin(p + 'mlp.fc2', D, Hd, std=0.02)
    ln('norm', D)

    sd['agg_block.latents'] = torch.from_numpy(rs.standard_normal(size=(num_latents, D)).astype(np.float32))
    inner = 4 * 512

    def agg_layer():
        t: State = {}
        t['0.fn.to_q.weight'] = _trunc_normal(rs, (inner, D), 0.03)
        t['0.fn.to_k.weight'] = _trunc_normal(rs, (inner, D), 0.03)
        t['0.fn.to_v.weight'] = _trunc_normal(rs, (inner, D), 0.03)
        t['0.fn.to_out.0.weight'] = _trunc_normal(rs, (D, inner), 0.02)
        t['0.fn.to_out.0.bias'] = _trunc_normal(rs, (D,), bias_std)
        t['0.norm.weight'] = 1.0 + _trunc_normal(rs, (D,), ln_jitter)
        t['0.norm.bias'] = _trunc_normal(rs, (D,), bias_std)
        t['0.norm_context.weight'] = 1.0 + _trunc_normal(rs, (D,), ln_jitter)
        t['0.norm_context.bias'] = _trunc_normal(rs, (D,), bias_std)
        t['2.fn.net.0.weight'] = _trunc_normal(rs, (4 * D, D), 0.03)
        t['2.fn.net.0.bias'] = _trunc_normal(rs, (4 * D,), bias_std)
        t['2.fn.net.3.weight'] = _trunc_normal(rs, (D, 4 * D), 0.02)
        t['2.fn.net.3.bias'] = _trunc_normal(rs, (D,), bias_std)
        t['2.norm.weight'] = 1.0 + _trunc_normal(rs, (D,), ln_jitter)
        t['2.norm.bias'] = _trunc_normal(rs, (D,), bias_std)
        return t

    shared = agg_layer() if agg_weights_tie else None
    for l in range(agg_depth):
        t = shared if agg_weights_tie else agg_layer()
        for k, v in t.items():
            sd[f'agg_block.layers.{l}.{k}'] = v
    ln('agg_block.last_layer.0', D)

    lin('mask_predictor.decoder.0', 512, D, std=0.03)
    lin('mask_predictor.decoder.2', 256, 512, std=0.04)
    lin('mask_predictor.decoder.4', 196, 256, std=0.05)
    lin('head', num_classes + num_scene_classes, D, std=head_std)
    return sd


def synth_teacher_state_dict(num_classes=365, depth=12, seed=0) -> State:
    """state_dict of the frozen scene teacher `vit_base_patch16_224(num_classes=365, use_mean_pooling=False)`
    (model/modeling_finetune.py:178-334): the student's encoder keys plus cls_token and a 365-way head."""
    full = synth_state_dict(num_classes=1, num_scene_classes=num_classes - 1, depth=depth, seed=seed + 500)
    sd = {k: v for k, v in full.items() if k.startswith(('patch_embed.', 'blocks.', 'norm.', 'head.'))}
    rs = np.random.RandomState(seed + 501)
    sd['cls_token'] = _trunc_normal(rs, (1, 1, 768), 0.02)
    return sd


def synth_clips(batch: int, seed=0, frames=16, size=224) -> Tensor:
    """N(0,1) clips [B,3,T,H,W] fp32 (ImageNet-normalised video is ~zero-mean/unit-var;
    dataset/kinetics.py:80-86).  numpy RandomState => identical on every machine."""
    rs = np.random.RandomState(1000 + seed)
    return torch.from_numpy(rs.standard_normal(size=(batch, 3, frames, size, size)).astype(np.float32))


def synth_tokens(batch: int, n_tokens=1568, dim=768, seed=0, scale=1.5) -> Tensor:
    """Slot micro-bench tokens: randn * 1.5."""
    rs = np.random.RandomState(2000 + seed)
    return torch.from_numpy((rs.standard_normal(size=(batch, n_tokens, dim)) * scale).astype(np.float32))


# ----------------------------------------------------------------------------------------------
# encoder (model/modeling_slot.py)
# ----------------------------------------------------------------------------------------------

def sinusoid_table(n_position: int, d_hid: int) -> Tensor:
    """model/modeling_slot.py:181-191 -- float64 numpy table, sin on even / cos on odd columns,
    angle = pos / 10000^(2*(j//2)/d_hid); returned as fp32 [1, n_position, d_hid]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.tensor(ang, dtype=torch.float).unsqueeze(0)


def patch_embed(sd: State, x: Tensor, tubelet=2, patch=16) -> Tensor:
    """model/modeling_slot.py:171-177 -- Conv3d(k=s=(2,16,16)) then flatten(2).transpose(1,2);
    restated as the equivalent GEMM over non-overlapping tubes (token = t*196 + h*14 + w,
    k = c*512 + dt*256 + dy*16 + dx)."""
    B, C, T, H, W = x.shape
    w = sd['patch_embed.proj.weight']
    D = w.shape[0]
    t, h, wv = T // tubelet, H // patch, W // patch
    cols = x.reshape(B, C, t, tubelet, h, patch, wv, patch).permute(0, 2, 4, 6, 1, 3, 5, 7)
    cols = cols.reshape(B, t * h * wv, C * tubelet * patch * patch)
    return cols @ w.reshape(D, -1).t() + sd['patch_embed.proj.bias']


def encoder_attention(sd: State, p: str, x: Tensor, num_heads=12) -> Tensor:
    """model/modeling_slot.py:95-117 -- qkv linear with bias cat(q_bias, 0, v_bias), q scaled
    BEFORE q.k^T, softmax over keys, proj."""
    B, N, C = x.shape
    qb, vb = sd[p + 'q_bias'], sd[p + 'v_bias']
    bias = torch.cat((qb, torch.zeros_like(vb), vb))
    qkv = F.linear(x, sd[p + 'qkv.weight'], bias).reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (q.shape[-1] ** -0.5)
    attn = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(x, sd[p + 'proj.weight'], sd[p + 'proj.bias'])


def encoder_mlp(sd: State, p: str, x: Tensor) -> Tensor:
    """model/modeling_slot.py:60-67 -- fc1, exact (erf) GELU, fc2."""
    x = F.gelu(F.linear(x, sd[p + 'fc1.weight'], sd[p + 'fc1.bias']))
    return F.linear(x, sd[p + 'fc2.weight'], sd[p + 'fc2.bias'])


def encoder_block(sd: State, i: int, x: Tensor, num_heads=12, eps=1e-6) -> Tensor:
    """model/modeling_slot.py:149-152 (return_attn=False branch; drop_path = identity in parity runs)."""
    p = f'blocks.{i}.'
    D = x.shape[-1]
    x = x + encoder_attention(sd, p + 'attn.', F.layer_norm(x, (D,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], eps), num_heads)
    x = x + encoder_mlp(sd, p + 'mlp.', F.layer_norm(x, (D,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], eps))
    return x


def forward_features(sd: State, clips: Tensor, depth: Optional[int] = None, num_heads=12, eps=1e-6) -> Tensor:
    """model/modeling_slot.py:350-377 -- patch embed, + sinusoid table, blocks, final LayerNorm(eps=1e-6)."""
    x = patch_embed(sd, clips)
    x = x + sinusoid_table(x.shape[1], x.shape[2]).type_as(x)
    if depth is None:
        depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('blocks.'))
    for i in range(depth):
        x = encoder_block(sd, i, x, num_heads, eps)
    D = x.shape[-1]
    return F.layer_norm(x, (D,), sd['norm.weight'], sd['norm.bias'], eps)


def teacher_forward(sd: State, clips: Tensor, depth: Optional[int] = None, num_heads=12, eps=1e-6):
    """model/modeling_finetune.py:270-325 with use_mean_pooling=False: CLS token prepended, 1569-row sinusoid table,
    blocks, LayerNorm, x[:, 0] -> head.  Returns (token [B,768], logits [B,365])."""
    x = patch_embed(sd, clips)
    x = torch.cat((sd['cls_token'].expand(x.shape[0], -1, -1), x), dim=1)
    x = x + sinusoid_table(x.shape[1], x.shape[2]).type_as(x)
    if depth is None:
        depth = 1 + max(int(k.split('.')[1]) for k in sd if k.startswith('blocks.'))
    for i in range(depth):
        x = encoder_block(sd, i, x, num_heads, eps)
    x = F.layer_norm(x, (x.shape[-1],), sd['norm.weight'], sd['norm.bias'], eps)
    token = x[:, 0]
    return token, F.linear(token, sd['head.weight'], sd['head.bias'])


# ----------------------------------------------------------------------------------------------
# aggregation block (agg_block/agg_block.py, agg_block/attention.py)
# ----------------------------------------------------------------------------------------------

def slot_cross_attention(sd: State, p: str, x: Tensor, ctx: Tensor, heads=4) -> Tuple[Tensor, Tensor]:
    """agg_block/attention.py:32-40 (PreNorm: LN(x), LN_ctx(context), eps 1e-5) +
    agg_block/attention.py:120-141: q/k/v projections without bias, 4 heads x 512,
    sim*512^-0.5, softmax over the SLOT axis (dim=1), sim_distill = that softmax,
    renormalise over tokens with +1e-7, attn.v, to_out."""
    D = x.shape[-1]
    xn = F.layer_norm(x, (D,), sd[p + 'norm.weight'], sd[p + 'norm.bias'], 1e-5)
    cn = F.layer_norm(ctx, (D,), sd[p + 'norm_context.weight'], sd[p + 'norm_context.bias'], 1e-5)
    q = F.linear(xn, sd[p + 'fn.to_q.weight'])
    k = F.linear(cn, sd[p + 'fn.to_k.weight'])
    v = F.linear(cn, sd[p + 'fn.to_v.weight'])
    B, S, inner = q.shape
    dh = inner // heads

    def split(t):  # 'b n (h d) -> (b h) n d'
        return t.reshape(B, t.shape[1], heads, dh).permute(0, 2, 1, 3).reshape(B * heads, t.shape[1], dh)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum('bid,bjd->bij', q, k) * (dh ** -0.5)
    attn = sim.softmax(dim=1)
    sim_distill = attn
    attn = attn / (attn.sum(dim=-1, keepdim=True) + 1e-7)
    out = torch.einsum('bij,bjd->bid', attn, v)
    out = out.reshape(B, heads, S, dh).permute(0, 2, 1, 3).reshape(B, S, inner)
    return F.linear(out, sd[p + 'fn.to_out.0.weight'], sd[p + 'fn.to_out.0.bias']), sim_distill


def slot_feed_forward(sd: State, p: str, x: Tensor) -> Tensor:
    """agg_block/attention.py:32-40 (PreNorm LN) + :63-69,81-82: Linear 768->3072, exact GELU, Linear 3072->768."""
    D = x.shape[-1]
    xn = F.layer_norm(x, (D,), sd[p + 'norm.weight'], sd[p + 'norm.bias'], 1e-5)
    h = F.gelu(F.linear(xn, sd[p + 'fn.net.0.weight'], sd[p + 'fn.net.0.bias']))
    return F.linear(h, sd[p + 'fn.net.3.weight'], sd[p + 'fn.net.3.bias'])


def aggregation_block(sd: State, data: Tensor, prefix='agg_block.', depth: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """agg_block/agg_block.py:120-139 -- slots start as the learned latents repeated over the batch;
    per layer: x = cross_attn(x, ctx) + x ; x = ff(x) + x ; finally LayerNorm (last_ln);
    returns (slots [B,S,D], sim of the LAST layer [(B*4),S,N])."""
    if depth is None:
        depth = 1 + max(int(k[len(prefix):].split('.')[1]) for k in sd if k.startswith(prefix + 'layers.'))
    B = data.shape[0]
    x = sd[prefix + 'latents'].unsqueeze(0).expand(B, -1, -1).type_as(data)
    sim = None
    for l in range(depth):
        lp = f'{prefix}layers.{l}.'
        a, sim = slot_cross_attention(sd, lp + '0.', x, data)
        x = a + x
        x = slot_feed_forward(sd, lp + '2.', x) + x
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), sd[prefix + 'last_layer.0.weight'], sd[prefix + 'last_layer.0.bias'], 1e-5)
    return x, sim


# ----------------------------------------------------------------------------------------------
# head / slot selection (model/modeling_slot.py:390-410) and the full student forward
# ----------------------------------------------------------------------------------------------

def mask_predictor(sd: State, slots: Tensor) -> Tensor:
    """model/modeling_slot.py:194-216 -- 768->512 ReLU ->256 ReLU ->196 Sigmoid."""
    h = F.relu(F.linear(slots, sd['mask_predictor.decoder.0.weight'], sd['mask_predictor.decoder.0.bias']))
    h = F.relu(F.linear(h, sd['mask_predictor.decoder.2.weight'], sd['mask_predictor.decoder.2.bias']))
    h = torch.sigmoid(F.linear(h, sd['mask_predictor.decoder.4.weight'], sd['mask_predictor.decoder.4.bias']))
    return h.squeeze().reshape(slots.shape[0], 196)


def head_matching(sd: State, slots: Tensor, attn: Tensor, num_classes: int, num_scene_classes=365):
    """model/modeling_slot.py:390-410 ('matching' branch; fc_dropout = identity in parity runs)."""
    bs, S, D = slots.shape
    flat = slots.reshape(-1, D)
    slots_head = F.linear(flat, sd['head.weight'], sd['head.bias'])
    probs = F.softmax(slots_head, dim=-1).view(bs, S, -1)
    a_idx = torch.argmax(probs[:, :, :num_classes].max(dim=-1).values, dim=1)
    s_idx = torch.argmax(probs[:, :, num_classes:num_classes + num_scene_classes].max(dim=-1).values, dim=1)
    ar = torch.arange(bs)
    action_feat = flat.view(bs, S, -1)[ar, a_idx]
    scene_feat = flat.view(bs, S, -1)[ar, s_idx]
    action_logit = slots_head.view(bs, S, -1)[ar, a_idx]
    scene_logit = slots_head.view(bs, S, -1)[ar, s_idx]
    mask_predictions = mask_predictor(sd, flat)
    return (action_feat, scene_feat), (action_logit, scene_logit, attn), (slots_head, flat, mask_predictions)


def student_forward(sd: State, clips: Tensor, num_classes: int, num_scene_classes=365,
                    depth: Optional[int] = None, agg_depth: Optional[int] = None):
    """model/modeling_slot.py:379-410 with slot_matching_method='matching'."""
    tokens = forward_features(sd, clips, depth)
    slots, attn = aggregation_block(sd, tokens, depth=agg_depth)
    return head_matching(sd, slots, attn, num_classes, num_scene_classes)



# ----------------------------------------------------------------------------------------------
# downstream fusion model (model/modeling_slot_fusion.py) -- SURVEY.md section 8f row N4
# ----------------------------------------------------------------------------------------------

def synth_fusion_state_dict(num_classes=101, num_scene_classes=365, num_latents=2, agg_depth=4, depth=2,
                            downstream_nb_classes=50, use_input_ln=True, seed=0) -> State:
    """state_dict of slot_fusion_vit_base_patch16_224(head_type='mlp', slot_fusion_method='concat'): the pre-training
    student's encoder / aggregation / head tensors plus action_norm, scene_norm and the MLPHead fusion_head
    (model/modeling_slot_fusion.py:23-38, 285-306).  No mask_predictor in this model."""
    sd = synth_state_dict(num_classes=num_classes, num_scene_classes=num_scene_classes, num_latents=num_latents,
                          agg_depth=agg_depth, agg_weights_tie=True, depth=depth, seed=seed)
    sd = {k: v for k, v in sd.items() if not k.startswith('mask_predictor.')}
    rs = np.random.RandomState(seed + 1000)
    D = 768

    def lin(prefix, out_f, in_f, std=0.03):
        sd[prefix + '.weight'] = _trunc_normal(rs, (out_f, in_f), std)
        sd[prefix + '.bias'] = _trunc_normal(rs, (out_f,), 0.02)

    def ln(prefix, dim):
        sd[prefix + '.weight'] = 1.0 + _trunc_normal(rs, (dim,), 0.1)
        sd[prefix + '.bias'] = _trunc_normal(rs, (dim,), 0.02)

    ln('action_norm', D); ln('scene_norm', D)
    lin('fusion_head.fc_action_down', D // 2, D); lin('fusion_head.fc_scene_down', D // 2, D)
    ln('fusion_head.fc_action_ln', D // 2); ln('fusion_head.fc_scene_ln', D // 2)
    if use_input_ln:
        ln('fusion_head.fc_input_ln', D)
    lin('fusion_head.classifier', downstream_nb_classes, D)
    return sd


def fusion_forward(sd: State, clips: Tensor, num_classes: int, num_scene_classes=365, depth: Optional[int] = None,
                   agg_depth: Optional[int] = None, use_input_ln=True, eps=1e-6):
    """model/modeling_slot_fusion.py:364-403 ('concat' + MLPHead :40-54; note the head's fc_action_* modules serve BOTH tokens)."""
    tokens = forward_features(sd, clips, depth)
    slots, _ = aggregation_block(sd, tokens, depth=agg_depth)
    bs, S, D = slots.shape
    flat = slots.reshape(-1, D)
    slots_head = F.linear(flat, sd['head.weight'], sd['head.bias'])
    probs = F.softmax(slots_head, dim=-1).view(bs, S, -1)
    a_idx = torch.argmax(probs[:, :, :num_classes].max(dim=-1).values, dim=1)
    s_idx = torch.argmax(probs[:, :, num_classes:num_classes + num_scene_classes].max(dim=-1).values, dim=1)
    ar = torch.arange(bs)
    a = F.layer_norm(flat.view(bs, S, -1)[ar, a_idx], (D,), sd['action_norm.weight'], sd['action_norm.bias'], eps)
    c = F.layer_norm(flat.view(bs, S, -1)[ar, s_idx], (D,), sd['scene_norm.weight'], sd['scene_norm.bias'], eps)
    inp = torch.cat((a, c), dim=1)

    def down(t):
        t = F.linear(t, sd['fusion_head.fc_action_down.weight'], sd['fusion_head.fc_action_down.bias'])
        return F.layer_norm(t, (D // 2,), sd['fusion_head.fc_action_ln.weight'], sd['fusion_head.fc_action_ln.bias'], 1e-5)

    out = torch.cat([down(a), down(c)], dim=1)
    if use_input_ln:
        out = F.layer_norm(out, (D,), sd['fusion_head.fc_input_ln.weight'], sd['fusion_head.fc_input_ln.bias'], 1e-5)
    out = F.linear(F.relu(out), sd['fusion_head.classifier.weight'], sd['fusion_head.classifier.bias'])
    return inp, out, (a_idx, s_idx)

# ----------------------------------------------------------------------------------------------
# training objective (utils/loss/train_loss.py:85-187) -- S=2.. small; Hungarian by brute force
# ----------------------------------------------------------------------------------------------

def _assign(cost: Tensor):
    """utils/loss/train_loss.py:121 scipy.linear_sum_assignment on an [S,2] cost: choose distinct
    slots (i for action, j for scene) minimising cost[i,0]+cost[j,1].  Brute force, ties -> the
    lexicographically first (i, j), which is what scipy's shortest-augmenting-path returns here."""
    S = cost.shape[0]
    best, arg = None, None
    c = cost.detach().double()
    for i in range(S):
        for j in range(S):
            if i == j:
                continue
            v = float(c[i, 0] + c[j, 1])
            if best is None or v < best:
                best, arg = v, (i, j)
    return arg


def train_loss(student_output, teacher_scene_logit: Tensor, target: Tensor, fg_mask, num_action_classes: int,
               scene_criterion='KL', scene_loss_weight=2000.0, mask_prediction_loss_weight=1.0,
               mask_distill_loss_weight=3.0):
    """utils/loss/train_loss.py:85-187 ('matching').  fp32 throughout (the reference's .half() on the
    masks, :136-137, is a dtype quirk of its fp16 recipe, not arithmetic)."""
    _, (action_output, _, attn), (slots_head, slots, mask_predictions) = student_output
    bs = target.shape[0]
    S = slots_head.shape[0] // bs
    H = attn.shape[0] // bs
    attn = attn.reshape(bs, H, S, -1).mean(dim=1)
    mask_predictions = mask_predictions.reshape(bs, S, -1)
    scene_target = torch.argmax(teacher_scene_logit, dim=1) + num_action_classes
    var = teacher_scene_logit.min() - 1.0
    teacher_full = torch.cat([torch.full((bs, num_action_classes), float(var), device=teacher_scene_logit.device), teacher_scene_logit], dim=1)
    sfm = slots_head.softmax(-1)
    slots_head = slots_head.view(bs, S, -1)
    fg, fg_frames = fg_mask
    action_loss = scene_loss = mp_loss = md_loss = 0.0
    action_logit = []
    for b in range(bs):
        cost = torch.stack([-sfm[b * S:(b + 1) * S, target[b]], -sfm[b * S:(b + 1) * S, scene_target[b]]], dim=1)
        i, j = _assign(cost)
        md_loss = md_loss + F.mse_loss(attn[b, i], fg_frames[b]) * mask_distill_loss_weight
        mp_loss = mp_loss + F.binary_cross_entropy_with_logits(mask_predictions[b, i], fg[b]) * mask_prediction_loss_weight
        action_loss = action_loss + F.cross_entropy(slots_head[b, i], target[b])
        action_logit.append(slots_head[b, i])
        if scene_criterion == 'CE':
            scene_loss = scene_loss + F.cross_entropy(slots_head[b, j], scene_target[b])
        else:
            scene_loss = scene_loss + F.kl_div(F.log_softmax(slots_head[b, j], dim=-1),
                                               F.log_softmax(teacher_full[b], dim=-1),
                                               reduction='batchmean', log_target=True) * scene_loss_weight
    action_loss, scene_loss, mp_loss, md_loss = (t / bs for t in (action_loss, scene_loss, mp_loss, md_loss))
    sl = F.normalize(slots.reshape(bs, S, -1), p=2, dim=2)
    cs = torch.bmm(sl, sl.transpose(1, 2)) * (1 - torch.eye(S, device=sl.device))
    cosine_loss = (cs.sum(dim=(1, 2)) / (S * (S - 1))).mean()
    total = action_loss + scene_loss + cosine_loss + mp_loss + md_loss
    parts = dict(action_loss=action_loss, scene_loss=scene_loss, cosine_loss=cosine_loss,
                 mask_prediction_loss=mp_loss, mask_distill_loss=md_loss)
    return total, torch.stack(action_logit), parts
