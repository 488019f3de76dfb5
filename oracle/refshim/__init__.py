"""Import the UNMODIFIED reference attention modules without mmcv (test infrastructure).

The reference's hot-path files import ``mmcv`` / ``mmdet`` at module scope and
those packages are absent from this image.  This shim
installs a minimal stand-in for exactly the names those files import --
registries, ``BaseModule``, ``xavier_init`` / ``constant_init``, identity
``force_fp32`` / ``auto_fp16`` decorators, and a ``TransformerLayerSequence`` --
and points ``mmcv.ops.multi_scale_deform_attn.multi_scale_deformable_attn_pytorch``
at the reference's own in-tree restatement
``multi_scale_deformable_attn_pytorch_2d`` (temporal_self_attention.py:293-348).
The reference source files are then executed from where they lie under
``/root/reference`` (nothing is copied).  Only ``tests/golden/make_golden.py`` and
the ``not gpu`` oracle-pinning tests use it, and only when ``/root/reference``
exists (it does not on the GPU box).
"""
import copy
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('APOLLO_REFERENCE_ROOT', '/root/reference')
_MOD_DIR = os.path.join(REFERENCE_ROOT, 'projects', 'mmdet3d_plugin', 'bevformer', 'modules')


def reference_available():
    return os.path.isfile(os.path.join(_MOD_DIR, 'spatial_cross_attention.py'))


class _Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, **default_args):
        return _build_from_cfg(cfg, self, default_args or None)


def _build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    typ = args.pop('type')
    cls = registry.get(typ) if isinstance(typ, str) else typ
    if cls is None:
        raise KeyError(f'{typ} is not in the {registry.name} registry')
    return cls(**args)


class _BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        pass


class _ConfigDict(dict):
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


def _identity_decorator_factory(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(fn):
        return fn
    return deco


def _deprecated_api_warning(name_dict, cls_name=None):
    def deco(fn):
        return fn
    return deco


def _xavier_init(module, gain=1, bias=0, distribution='normal'):
    if module is None:
        return
    if hasattr(module, 'weight') and module.weight is not None:
        if distribution == 'uniform':
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _constant_init(module, val, bias=0):
    if hasattr(module, 'weight') and module.weight is not None:
        nn.init.constant_(module.weight, val)
    if hasattr(module, 'bias') and module.bias is not None:
        nn.init.constant_(module.bias, bias)


def _digit_version(v):
    out = []
    for tok in str(v).split('+')[0].split('.'):
        num = ''.join(ch for ch in tok if ch.isdigit())
        out.append(int(num) if num else 0)
    return tuple(out)


class _ExtLoader:
    @staticmethod
    def load_ext(name, funcs):
        ext = types.SimpleNamespace()

        def _missing(*a, **k):
            raise RuntimeError('mmcv _ext is not available in the reference shim (CPU path only)')
        for f in funcs:
            setattr(ext, f, _missing)
        return ext


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load_file(modname, path):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_loaded = None


def load_reference():
    """Returns a namespace with the reference's own classes / functions (CPU branch)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise FileNotFoundError(f'reference sources not found under {REFERENCE_ROOT}')
    saved = {k: v for k, v in sys.modules.items()
             if k.split('.')[0] in ('mmcv', 'mmdet', 'projects', 'matplotlib')}

    ATTENTION = _Registry('attention')
    FFN_REG = _Registry('feed-forward network')
    POS_REG = _Registry('position encoding')
    LAYER = _Registry('transformerLayer')
    LAYER_SEQ = _Registry('transformer-layers sequence')

    class TransformerLayerSequence(_BaseModule):
        def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
            super().__init__(init_cfg)
            if isinstance(transformerlayers, dict):
                transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
            self.num_layers = num_layers
            self.layers = nn.ModuleList()
            for cfg in (transformerlayers or []):
                self.layers.append(_build_from_cfg(cfg, LAYER))

    def build_attention(cfg, default_args=None):
        return _build_from_cfg(cfg, ATTENTION, default_args)

    tsa_path = os.path.join(_MOD_DIR, 'temporal_self_attention.py')

    def _lazy_msda_pytorch(*args, **kwargs):
        tsa = sys.modules['projects.mmdet3d_plugin.bevformer.modules.temporal_self_attention']
        return tsa.multi_scale_deformable_attn_pytorch_2d(*args, **kwargs)

    mmcv = _module('mmcv', ConfigDict=_ConfigDict, deprecated_api_warning=_deprecated_api_warning)
    mmcv.__path__ = []
    _module('mmcv.ops').__path__ = []
    _module('mmcv.ops.multi_scale_deform_attn',
            multi_scale_deformable_attn_pytorch=_lazy_msda_pytorch)
    cnn = _module('mmcv.cnn', xavier_init=_xavier_init, constant_init=_constant_init,
                  Linear=nn.Linear)
    cnn.__path__ = []
    _module('mmcv.cnn.bricks').__path__ = []
    _module('mmcv.cnn.bricks.registry', ATTENTION=ATTENTION, TRANSFORMER_LAYER=LAYER,
            TRANSFORMER_LAYER_SEQUENCE=LAYER_SEQ, FEEDFORWARD_NETWORK=FFN_REG,
            POSITIONAL_ENCODING=POS_REG)
    def build_transformer_layer_sequence(cfg, default_args=None):
        return _build_from_cfg(cfg, LAYER_SEQ, default_args)

    # mmcv 1.4.0 bricks the MapTRv2 decoder layer derives from / is configured with, restated from
    # their documented behaviour (third-party code, not under /root/reference): MultiheadAttention
    # (nn.MultiheadAttention + positional adds + identity), FFN, BaseTransformerLayer.__init__.
    class MultiheadAttention(_BaseModule):
        def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0.,
                     dropout_layer=dict(type='Dropout', drop_prob=0.), init_cfg=None,
                     batch_first=False, **kwargs):
            super().__init__(init_cfg)
            if 'dropout' in kwargs:
                attn_drop = kwargs.pop('dropout')
                dropout_layer = dict(type='Dropout', drop_prob=attn_drop)
            self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
            self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
            self.proj_drop = nn.Dropout(proj_drop)
            p = (dropout_layer or {}).get('drop_prob', 0.)
            self.dropout_layer = nn.Dropout(p) if dropout_layer else nn.Identity()

        def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                    attn_mask=None, key_padding_mask=None, **kwargs):
            if key is None:
                key = query
            if value is None:
                value = key
            if identity is None:
                identity = query
            if key_pos is None and query_pos is not None and query_pos.shape == key.shape:
                key_pos = query_pos
            if query_pos is not None:
                query = query + query_pos
            if key_pos is not None:
                key = key + key_pos
            if self.batch_first:
                query, key, value = (t.transpose(0, 1) for t in (query, key, value))
            out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                            key_padding_mask=key_padding_mask)[0]
            if self.batch_first:
                out = out.transpose(0, 1)
            return identity + self.dropout_layer(self.proj_drop(out))

    ATTENTION.module_dict['MultiheadAttention'] = MultiheadAttention

    class FFN(_BaseModule):
        def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                     act_cfg=dict(type='ReLU', inplace=True), ffn_drop=0., dropout_layer=None,
                     add_identity=True, init_cfg=None, **kwargs):
            super().__init__(init_cfg)
            layers, cin = [], embed_dims
            for _ in range(num_fcs - 1):
                layers.append(nn.Sequential(nn.Linear(cin, feedforward_channels), nn.ReLU(inplace=True),
                                            nn.Dropout(ffn_drop)))
                cin = feedforward_channels
            layers.append(nn.Linear(feedforward_channels, embed_dims))
            layers.append(nn.Dropout(ffn_drop))
            self.layers = nn.Sequential(*layers)
            self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()
            self.add_identity = add_identity

        def forward(self, x, identity=None):
            out = self.layers(x)
            if not self.add_identity:
                return self.dropout_layer(out)
            if identity is None:
                identity = x
            return identity + self.dropout_layer(out)

    class BaseTransformerLayer(_BaseModule):
        def __init__(self, attn_cfgs=None, ffn_cfgs=None, operation_order=None, norm_cfg=dict(type='LN'),
                     init_cfg=None, batch_first=False, **kwargs):
            super().__init__(init_cfg)
            ffn_cfgs = dict(ffn_cfgs or dict(type='FFN', embed_dims=256, feedforward_channels=1024,
                                             num_fcs=2, ffn_drop=0.))
            for old_name, new_name in (('feedforward_channels', 'feedforward_channels'),
                                       ('ffn_dropout', 'ffn_drop'), ('ffn_num_fcs', 'num_fcs')):
                if old_name in kwargs:
                    ffn_cfgs[new_name] = kwargs[old_name]
            self.batch_first = batch_first
            num_attn = operation_order.count('self_attn') + operation_order.count('cross_attn')
            if isinstance(attn_cfgs, dict):
                attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
            assert num_attn == len(attn_cfgs)
            self.num_attn = num_attn
            self.operation_order = operation_order
            self.norm_cfg = norm_cfg
            self.pre_norm = operation_order[0] == 'norm'
            self.attentions = nn.ModuleList()
            index = 0
            for name in operation_order:
                if name in ('self_attn', 'cross_attn'):
                    cfg = copy.deepcopy(attn_cfgs[index])
                    cfg['batch_first'] = batch_first
                    att = build_attention(cfg)
                    att.operation_name = name
                    self.attentions.append(att)
                    index += 1
            self.embed_dims = self.attentions[0].embed_dims
            ffn_cfgs.pop('type', None)
            ffn_cfgs['embed_dims'] = self.embed_dims
            self.ffns = nn.ModuleList([FFN(**copy.deepcopy(ffn_cfgs))
                                       for _ in range(operation_order.count('ffn'))])
            self.norms = nn.ModuleList([nn.LayerNorm(self.embed_dims)
                                        for _ in range(operation_order.count('norm'))])

    def inverse_sigmoid(x, eps=1e-5):              # mmdet.models.utils.transformer
        x = x.clamp(min=0, max=1)
        return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))

    FFN_REG.module_dict['FFN'] = FFN

    def build_feedforward_network(cfg, default_args=None):
        return _build_from_cfg(cfg, FFN_REG, default_args)

    def build_norm_layer(cfg, num_features, postfix=''):
        assert cfg.get('type') == 'LN', cfg
        return 'ln' + str(postfix), nn.LayerNorm(num_features, eps=cfg.get('eps', 1e-5))

    def build_activation_layer(cfg):
        assert cfg.get('type') == 'ReLU', cfg
        return nn.ReLU(inplace=cfg.get('inplace', False))

    cnn.build_norm_layer = build_norm_layer
    cnn.build_activation_layer = build_activation_layer
    _module('mmcv.cnn.bricks.transformer', build_attention=build_attention,
            TransformerLayerSequence=TransformerLayerSequence, BaseTransformerLayer=BaseTransformerLayer,
            build_feedforward_network=build_feedforward_network,
            build_transformer_layer_sequence=build_transformer_layer_sequence)
    TRANSFORMER = _Registry('Transformer')
    for pkg in ('mmdet', 'mmdet.models', 'mmdet.models.utils'):
        _module(pkg).__path__ = []
    _module('mmdet.models.utils.builder', TRANSFORMER=TRANSFORMER)
    _module('mmdet.models.utils.transformer', inverse_sigmoid=inverse_sigmoid)
    runner = _module('mmcv.runner', force_fp32=_identity_decorator_factory,
                     auto_fp16=_identity_decorator_factory)
    runner.__path__ = []
    _module('mmcv.runner.base_module', BaseModule=_BaseModule, ModuleList=nn.ModuleList,
            Sequential=nn.Sequential)
    _module('mmcv.utils', ext_loader=_ExtLoader, ConfigDict=_ConfigDict,
            build_from_cfg=_build_from_cfg, deprecated_api_warning=_deprecated_api_warning,
            to_2tuple=lambda x: (x, x), TORCH_VERSION=torch.__version__,
            digit_version=_digit_version)
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            _module('matplotlib').__path__ = []
            _module('matplotlib.pyplot')

    for pkg in ('projects', 'projects.mmdet3d_plugin', 'projects.mmdet3d_plugin.bevformer',
                'projects.mmdet3d_plugin.bevformer.modules', 'projects.mmdet3d_plugin.models',
                'projects.mmdet3d_plugin.models.utils'):
        _module(pkg).__path__ = []
    _module('projects.mmdet3d_plugin.models.utils.bricks',
            run_time=lambda name: (lambda fn: fn))
    _module('projects.mmdet3d_plugin.models.utils.visual', save_tensor=lambda *a, **k: None)
    prefix = 'projects.mmdet3d_plugin.bevformer.modules.'
    try:
        # the reference's own base layer (attention / FFN / norm construction and the generic
        # operation loop BEVFormerLayer falls back to), executed from the reference tree
        _load_file(prefix + 'custom_base_transformer_layer',
                   os.path.join(_MOD_DIR, 'custom_base_transformer_layer.py'))
        fn_mod = _load_file(prefix + 'multi_scale_deformable_attn_function',
                            os.path.join(_MOD_DIR, 'multi_scale_deformable_attn_function.py'))
        tsa = _load_file(prefix + 'temporal_self_attention', tsa_path)
        sca = _load_file(prefix + 'spatial_cross_attention',
                         os.path.join(_MOD_DIR, 'spatial_cross_attention.py'))
        enc = _load_file(prefix + 'encoder', os.path.join(_MOD_DIR, 'encoder.py'))
        dec = _load_file(prefix + 'decoder', os.path.join(_MOD_DIR, 'decoder.py'))
        try:                      # needs torchvision (prev_bev rotation); optional
            trf = _load_file(prefix + 'transformer', os.path.join(_MOD_DIR, 'transformer.py'))
        except ImportError:
            trf = None
        mapdec = _load_file('projects.mmdet3d_plugin.maptrv2.modules.decoder',
                            os.path.join(REFERENCE_ROOT, 'projects', 'mmdet3d_plugin', 'maptrv2',
                                         'modules', 'decoder.py'))
    finally:
        # keep the synthetic entries only as long as needed by the loaded modules' globals
        for k in list(sys.modules):
            if k.split('.')[0] in ('mmcv', 'mmdet', 'projects') and k not in saved:
                if not k.startswith(prefix):
                    del sys.modules[k]
        sys.modules.update(saved)

    _loaded = types.SimpleNamespace(
        msda_pytorch_2d=tsa.multi_scale_deformable_attn_pytorch_2d,
        TemporalSelfAttention=tsa.TemporalSelfAttention,
        SpatialCrossAttention=sca.SpatialCrossAttention,
        MSDeformableAttention3D=sca.MSDeformableAttention3D,
        BEVFormerEncoder=enc.BEVFormerEncoder,
        BEVFormerLayer=enc.BEVFormerLayer,
        CustomMSDeformableAttention=dec.CustomMSDeformableAttention,
        DetectionTransformerDecoder=dec.DetectionTransformerDecoder,
        MapTRv2Decoder=mapdec.MapTRv2Decoder,
        MapTRv2DecoupledDetrTransformerDecoderLayer=mapdec.MapTRv2DecoupledDetrTransformerDecoderLayer,
        Function_fp32=fn_mod.MultiScaleDeformableAttnFunction_fp32,
        Function_fp16=fn_mod.MultiScaleDeformableAttnFunction_fp16,
        PerceptionTransformer=None if trf is None else trf.PerceptionTransformer,
        ATTENTION=ATTENTION, LAYER=LAYER, LAYER_SEQ=LAYER_SEQ, build_attention=build_attention,
        modules=dict(tsa=tsa, sca=sca, enc=enc, dec=dec, fn=fn_mod, trf=trf, mapdec=mapdec))
    return _loaded


def load_reference_dcnv3():
    """The reference's ``ops_dcnv3/functions/dcnv3_func.py`` (unmodified, executed from /root/reference) with a
    stand-in for its compiled ``DCNv3`` extension, which cannot be built here: only the pure-PyTorch
    ``dcnv3_core_pytorch`` (:119-188, "for debug and test only") is usable, and that is what the DCNv3 golden
    vectors are generated with."""
    import importlib.util
    import types
    path = os.path.join(REFERENCE_ROOT, 'projects', 'mmdet3d_plugin', 'bevformer', 'backbones', 'ops_dcnv3',
                        'functions', 'dcnv3_func.py')
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    saved = sys.modules.get('DCNv3')
    sys.modules['DCNv3'] = types.ModuleType('DCNv3')
    try:
        spec = importlib.util.spec_from_file_location('_reference_dcnv3_func', path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop('DCNv3', None)
        else:
            sys.modules['DCNv3'] = saved
    return mod
