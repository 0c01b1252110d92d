"""Import the *unmodified* reference hot path from /root/reference (TEST INFRASTRUCTURE).

The reference needs `timm` for exactly four symbols (model/modeling_slot.py:6-7); timm is not
installed and there is no network, so an in-memory shim provides them.  Nothing from the reference
is copied: the modules are imported from where they lie.  /root/reference only exists in the build
container -- callers must check `available()` first; on the GPU box the committed golden vectors
(tests/golden/) stand in for it.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get('DEVIAS_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'model', 'modeling_slot.py'))


def _install_timm_shim():
    if 'timm' in sys.modules and not getattr(sys.modules['timm'], '_devias_shim', False):
        return  # a real timm is importable: use it
    timm = types.ModuleType('timm'); timm._devias_shim = True
    models = types.ModuleType('timm.models')
    layers = types.ModuleType('timm.models.layers')
    registry = types.ModuleType('timm.models.registry')

    def drop_path(x, drop_prob: float = 0., training: bool = False):
        # timm 0.4.12 semantics: per-sample Bernoulli(keep) / keep
        if drop_prob == 0. or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = (keep + torch.rand(shape, dtype=x.dtype, device=x.device)).floor_()
        return x.div(keep) * mask

    def to_2tuple(v):
        return tuple(v) if isinstance(v, (tuple, list)) else (v, v)

    layers.drop_path = drop_path
    layers.to_2tuple = to_2tuple
    layers.trunc_normal_ = torch.nn.init.trunc_normal_
    registry.register_model = lambda fn: fn
    timm.models = models; models.layers = layers; models.registry = registry
    sys.modules.update({'timm': timm, 'timm.models': models, 'timm.models.layers': layers,
                        'timm.models.registry': registry})


def load():
    """Returns a namespace with the reference modules: .modeling_slot, .agg_block, .attention, .train_loss."""
    if not available():
        raise RuntimeError(f'reference tree not present at {REFERENCE_ROOT}')
    try:
        import timm  # noqa: F401
    except Exception:
        _install_timm_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    with contextlib.redirect_stdout(io.StringIO()):
        import importlib
        ns.agg_block = importlib.import_module('agg_block.agg_block')
        ns.attention = importlib.import_module('agg_block.attention')
        ns.modeling_slot = importlib.import_module('model.modeling_slot')
        ns.train_loss = importlib.import_module('utils.loss.train_loss')
    return ns


def build_student(ns, **kwargs):
    """slot_vit_base_patch16_224 with parity-friendly defaults."""
    kw = dict(num_classes=101, all_frames=16, tubelet_size=2, drop_path_rate=0., fc_drop_rate=0.,
              init_scale=1.0, num_latents=2, head_type='linear', slot_matching_method='matching',
              agg_weights_tie=True, agg_depth=4, num_scene_classes=365)
    kw.update(kwargs)
    with contextlib.redirect_stdout(io.StringIO()):
        if 'depth' in kw:  # reduced-depth fixtures: same fixed ctor args as modeling_slot.py:417-420
            from functools import partial
            m = ns.modeling_slot.VisionTransformer(
                patch_size=16, embed_dim=768, num_heads=12, mlp_ratio=4, qkv_bias=True,
                norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **kw)
        else:
            m = ns.modeling_slot.slot_vit_base_patch16_224(pretrained=False, **kw)
    return m
