"""Plugin registry for the attention modules.

The reference builds its modules through mmcv's registries from ``dict(type=...)`` config
fragments (``@ATTENTION.register_module()``, ``build_attention(cfg)``;
spatial_cross_attention.py:28,61, custom_base_transformer_layer.py:133).  When mmcv is
importable the B200 modules register into mmcv's own registries, so the reference's configs
build them unchanged; otherwise an equivalent local registry with the same
``register_module`` / ``build_from_cfg`` semantics is used (mmcv is absent from this image).
"""
import copy

import torch.nn as nn

try:  # pragma: no cover - mmcv is not installed in the build image
    from mmcv.cnn.bricks.registry import (ATTENTION, FEEDFORWARD_NETWORK, TRANSFORMER_LAYER,
                                          TRANSFORMER_LAYER_SEQUENCE)
    from mmcv.runner.base_module import BaseModule
    from mmcv.utils import build_from_cfg
    HAVE_MMCV = True
except Exception:
    HAVE_MMCV = False

    class Registry:
        def __init__(self, name):
            self.name = name
            self.module_dict = {}

        def register_module(self, name=None, force=False, module=None):
            def deco(cls):
                key = name or cls.__name__
                if key in self.module_dict and not force:
                    raise KeyError(f'{key} is already registered in {self.name}')
                self.module_dict[key] = cls
                return cls
            return deco(module) if module is not None else deco

        def get(self, key):
            return self.module_dict.get(key)

        def build(self, cfg, **default_args):
            return build_from_cfg(cfg, self, default_args or None)

    def build_from_cfg(cfg, registry, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise KeyError(f'cfg must be a dict with a "type" key, got {cfg}')
        args = copy.deepcopy(dict(cfg))
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        typ = args.pop('type')
        cls = registry.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f'{typ} is not in the {registry.name} registry')
        return cls(**args)

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self._is_init = False
            self.init_cfg = copy.deepcopy(init_cfg)

        def init_weights(self):
            for m in self.children():
                if hasattr(m, 'init_weights'):
                    m.init_weights()

    ATTENTION = Registry('attention')
    FEEDFORWARD_NETWORK = Registry('feed-forward network')
    TRANSFORMER_LAYER = Registry('transformerLayer')
    TRANSFORMER_LAYER_SEQUENCE = Registry('transformer-layers sequence')


try:  # pragma: no cover - mmdet is not installed in the build image
    from mmdet.models.utils.builder import TRANSFORMER
except Exception:
    if HAVE_MMCV:  # pragma: no cover
        from mmcv.utils import Registry as _MMCVRegistry
        TRANSFORMER = _MMCVRegistry('Transformer')
    else:
        TRANSFORMER = Registry('Transformer')


def build_transformer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER, default_args)


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


def xavier_init(module, gain=1, bias=0, distribution='normal'):
    if module is None:
        return
    if getattr(module, 'weight', None) is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)


def constant_init(module, val, bias=0):
    if getattr(module, 'weight', None) is not None:
        nn.init.constant_(module.weight, val)
    if getattr(module, 'bias', None) is not None:
        nn.init.constant_(module.bias, bias)
