"""apollo-vision-net_b200: B200-native (sm_100a) multi-scale deformable attention for the
BEVFormer encoder / MapTRv2 decoder hot path of HankerSia/Apollo-Vision-Net.

Import it as ``apollo_vision_net_b200`` (``apollo_vision_net_b200.py`` at the repo root maps
that name onto this directory, whose name is not a Python identifier).

Nothing here imports ``oracle/``; every operator runs through ``libmsda_b200.so`` and raises
when that library is missing.
"""
from . import _lib
from ._lib import build, launch_count
from . import bev_prep, dcnv3, fused_ops, mha, modules, parallel, registry, rowops, synthetic
from .fused_ops import (QueueDeformAttnFunction, SpatialCrossAttnFunction, bev_point_sampling)
from .registry import (ATTENTION, TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE, build_attention,
                       build_transformer_layer, build_transformer_layer_sequence)
from .multi_scale_deformable_attn_function import (
    MultiScaleDeformableAttnFunction_bf16, MultiScaleDeformableAttnFunction_fp16,
    MultiScaleDeformableAttnFunction_fp32, ext_module,
    ms_deform_attn_backward, ms_deform_attn_forward)

__all__ = ['build', 'launch_count', 'ext_module', 'ms_deform_attn_forward',
           'ms_deform_attn_backward', 'MultiScaleDeformableAttnFunction_fp32',
           'MultiScaleDeformableAttnFunction_fp16', 'MultiScaleDeformableAttnFunction_bf16']
