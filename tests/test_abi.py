"""CPU: the C-ABI shared library loads and exports every symbol include/devias_b200.h declares (no compute calls),
the ctypes binding table covers all of them, and the product path refuses to run without CUDA."""
import contextlib
import ctypes
import io
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'devias_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(devias_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from devias_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    names = _declared()
    assert len(names) >= 12
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/devias_b200.h but not exported'
    assert sorted(_lib.exported_symbols()) == names, 'ctypes binding table and header disagree'
    l = _lib.lib()
    assert l.devias_abi_version() >= 1
    assert l.devias_launch_count() >= 0


def test_bad_arguments_return_status_not_crash():
    from devias_b200 import _lib
    l = _lib.lib()
    rc = l.devias_gemm_bf16(None, 0, 0, None, 0, 0, 0, 0, 0, 0, None, 0, None, 0, None, None, 0, 0, None, 0, 1, None)
    assert rc == 1 and b'null operand' in l.devias_last_error()
    rc = l.devias_layernorm_fwd(None, None, None, None, 1, None, None, 4, 768, 1e-5, None)
    assert rc == 1


def test_no_cpu_fallback():
    from devias_b200 import ops
    from devias_b200.agg_block import AggregationBlock
    with pytest.raises(RuntimeError):
        ops.cast_bf16(torch.zeros(16))
    with contextlib.redirect_stdout(io.StringIO()):
        m = AggregationBlock(num_latents=2, depth=1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1568, 768))
    with pytest.raises(AssertionError):
        m(torch.zeros(1, 2, 1568, 768))


def test_model_interface_matches_reference_contract():
    """constructor kwargs / attributes / state_dict keys the reference's scripts rely on"""
    from devias_b200.modeling_slot import slot_vit_base_patch16_224
    from oracle import devias_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        m = slot_vit_base_patch16_224(num_classes=400, all_frames=16, tubelet_size=2, fc_drop_rate=0.0, drop_rate=0.0,
                                      drop_path_rate=0.1, attn_drop_rate=0.0, use_checkpoint=False, init_scale=0.001,
                                      num_latents=2, head_type='linear', slot_matching='matching',
                                      slot_matching_method='matching', agg_weights_tie=True, agg_depth=8, num_scene_classes=365)
    sd = O.synth_state_dict(num_classes=400, num_latents=2, agg_depth=8, agg_weights_tie=True)
    assert set(m.state_dict().keys()) == set(sd.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    assert sum(p.numel() for p in m.parameters()) == 98413249
    assert 'pos_embed' not in m.state_dict() and m.pos_embed.shape == (1, 1568, 768)
    assert m.get_num_layers() == 12 and m.no_weight_decay() == {'pos_embed', 'cls_token'}
    assert m.patch_embed.patch_size == (16, 16) and m.patch_embed.num_patches == 1568 and m.patch_embed.tubelet_size == 2
    assert m.default_cfg['num_classes'] == 400
    assert m.agg_block.layers[0][0] is m.agg_block.layers[7][0]      # tied layers alias one module (attention.py:12-23)
    assert abs(m.blocks[11].drop_path.drop_prob - 0.1) < 1e-6 and isinstance(m.blocks[0].drop_path, torch.nn.Identity)
    with pytest.raises(ValueError):
        with contextlib.redirect_stdout(io.StringIO()):
            slot_vit_base_patch16_224(num_classes=10, slot_matching_method='bogus')


def test_reference_checkpoint_loader_walk_and_layer_decay_names():
    """INTEGRATION.md claims: (1) the reference's own checkpoint loader (utils/utils.py:330-348: a recursive walk over
    `_modules` calling `_load_from_state_dict` with the running prefix) fills the drop-in model completely from a state_dict with
    the reference's keys -- a VideoMAE pre-training checkpoint (encoder keys only) leaves exactly the slot / head parameters
    missing; (2) parameter NAMES map to the layer-decay groups of utils/optim_factory.py:24-35."""
    from functools import partial
    from devias_b200.modeling_slot import VisionTransformer
    from oracle import devias_oracle as O
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisionTransformer(patch_size=16, embed_dim=768, depth=2, num_heads=12, mlp_ratio=4, qkv_bias=True,
                              norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_classes=11, num_latents=2, agg_depth=2,
                              agg_weights_tie=True, slot_matching_method='matching', init_scale=1.0)
    sd = O.synth_state_dict(num_classes=11, num_latents=2, agg_depth=2, depth=2, seed=5)

    def walk_load(model, state_dict):          # the loader of utils/utils.py:330-348, restated
        missing, unexpected, errors = [], [], []
        state_dict = dict(state_dict)

        def load(module, prefix=''):
            module._load_from_state_dict(state_dict, prefix, {}, True, missing, unexpected, errors)
            for name, child in module._modules.items():
                if child is not None:
                    load(child, prefix + name + '.')
        load(model)
        return missing, errors

    missing, errors = walk_load(m, sd)
    assert not missing and not errors
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    encoder_only = {k: v for k, v in sd.items() if k.startswith(('patch_embed.', 'blocks.', 'norm.'))}
    missing, errors = walk_load(m, encoder_only)
    assert not errors and missing and all(k.startswith(('agg_block.', 'mask_predictor.', 'head.')) for k in missing)

    def layer_id(name, num_layers):            # utils/optim_factory.py:24-35 (get_num_layer_for_vit), restated
        if name in ('cls_token', 'mask_token', 'pos_embed') or name.startswith('patch_embed'):
            return 0
        if name.startswith('rel_pos_bias'):
            return num_layers - 1
        if name.startswith('blocks'):
            return int(name.split('.')[1]) + 1
        return num_layers - 1
    n = m.get_num_layers() + 2
    ids = {k: layer_id(k, n) for k, _ in m.named_parameters()}
    assert ids['patch_embed.proj.weight'] == 0 and ids['blocks.1.mlp.fc1.weight'] == 2 and ids['head.weight'] == n - 1
    assert any('agg_block' in k for k in ids)          # the 'agg_block' substring selects agg_block_scale (optim_factory.py:66-78)
