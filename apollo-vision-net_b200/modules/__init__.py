"""Registry-compatible attention modules of the hot path (same names as the reference's
``projects/mmdet3d_plugin/bevformer/modules`` and ``maptrv2/modules``)."""
from .spatial_cross_attention import MSDeformableAttention3D, SpatialCrossAttention
from .temporal_self_attention import TemporalSelfAttention
from .decoder import CustomMSDeformableAttention, inverse_sigmoid
from .encoder import FFN, BEVFormerEncoder, BEVFormerLayer
from .maptrv2_decoder import MapTRv2DecoupledDetrTransformerDecoderLayer, MapTRv2Decoder
from .detection_decoder import DetectionTransformerDecoder
from .transformer import PerceptionTransformer

__all__ = ['SpatialCrossAttention', 'MSDeformableAttention3D', 'TemporalSelfAttention',
           'CustomMSDeformableAttention', 'inverse_sigmoid', 'BEVFormerEncoder', 'BEVFormerLayer',
           'FFN', 'DetectionTransformerDecoder', 'MapTRv2Decoder', 'MapTRv2DecoupledDetrTransformerDecoderLayer', 'PerceptionTransformer']
