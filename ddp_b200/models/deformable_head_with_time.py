"""``DeformableHeadWithTime`` — plug-in surface of the reference decode head.

Reference: segmentation/mmseg/models/decode_heads/deformable_head_with_time.py:21-189 (seg),
depth/depth/models/decode_heads/deformable_head_with_time.py:20-131 (depth).  The class keeps the reference's
constructor signature, attributes (in_channels, channels, num_classes, out_channels, align_corners, ...) and
state-dict keys; its parameters are plain torch parameters so that reference checkpoints load with
``load_state_dict``.  The arithmetic runs in libddp_b200.so: the owning segmentor drives the whole T-step
loop through ``DecodeEngine``; ``forward(inputs, times)`` (one denoiser call) is served by the same library.
"""
import math

import torch
import torch.nn as nn

from ..registry import HEADS

EMBED, HEADS_N, POINTS, LEVELS = 256, 8, 4, 1


class _MSDeformAttnParams(nn.Module):
    """Parameters of mmcv's MultiScaleDeformableAttention (mmcv/ops/multi_scale_deform_attn.py:222-248)."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=4, dropout=0.1, **kw):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError(f"embed_dims must be divisible by num_heads, but got {embed_dims} and {num_heads}")
        if (embed_dims, num_heads, num_levels, num_points) != (EMBED, HEADS_N, LEVELS, POINTS):
            raise NotImplementedError("libddp_b200 is built for embed_dims=256, num_heads=8, num_levels=1, num_points=4 "
                                      "(every shipped DDP config)")
        self.embed_dims, self.num_heads, self.num_levels, self.num_points = embed_dims, num_heads, num_levels, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid.view(-1)
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)


class _FFNParams(nn.Module):
    """mmcv FFN (cnn/bricks/transformer.py:253-263): layers = [Sequential(Linear, act, Dropout), Linear, Dropout]."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, ffn_drop=0., act_cfg=None, **kw):
        super().__init__()
        if num_fcs != 2 or embed_dims != EMBED or feedforward_channels != 1024:
            raise NotImplementedError("libddp_b200 is built for FFN(256 -> 1024 -> 256)")
        if (act_cfg or {}).get("type", "ReLU") != "GELU":
            raise NotImplementedError("libddp_b200 implements the GELU FFN of the DDP configs")
        self.layers = nn.Sequential(nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.GELU(), nn.Dropout(ffn_drop)),
                                    nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))


class _TimeLayerParams(nn.Module):
    """The reference's forked BaseTransformerLayer (segmentation/mmseg/models/utils/transformer.py:182-316)."""

    def __init__(self, attn_cfgs, ffn_cfgs, operation_order, use_time_mlp=False, norm_cfg=None, **kw):
        super().__init__()
        if tuple(operation_order) != ("self_attn", "norm", "ffn", "norm"):
            raise NotImplementedError("libddp_b200 implements operation_order=('self_attn','norm','ffn','norm')")
        if not use_time_mlp:
            raise NotImplementedError("libddp_b200 implements the time-conditioned layer (use_time_mlp=True)")
        if (norm_cfg or dict(type="LN")).get("type") != "LN":
            raise NotImplementedError("libddp_b200 implements LayerNorm layers")
        attn = dict(attn_cfgs)
        if attn.pop("type") != "MultiScaleDeformableAttention":
            raise NotImplementedError("attention must be MultiScaleDeformableAttention")
        ffn = dict(ffn_cfgs)
        ffn.pop("type", None)
        self.attentions = nn.ModuleList([_MSDeformAttnParams(**attn)])
        self.ffns = nn.ModuleList([_FFNParams(**ffn)])
        self.norms = nn.ModuleList([nn.LayerNorm(EMBED), nn.LayerNorm(EMBED)])
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(EMBED * 4, EMBED * 2))
        self.embed_dims = EMBED


class _EncoderParams(nn.Module):
    """DetrTransformerEncoder / TransformerLayerSequence: ``layers`` ModuleList (transformer.py:1300-1329)."""

    def __init__(self, transformerlayers, num_layers, type="DetrTransformerEncoder", post_norm_cfg=None, **kw):
        super().__init__()
        if type != "DetrTransformerEncoder":
            raise NotImplementedError(f"encoder type {type}")
        cfg = dict(transformerlayers)
        if cfg.pop("type") != "BaseTransformerLayer":
            raise NotImplementedError("transformerlayers.type must be BaseTransformerLayer")
        if not 1 <= num_layers <= 8:
            raise NotImplementedError("libddp_b200 supports 1..8 encoder layers")
        self.layers = nn.ModuleList([_TimeLayerParams(**cfg) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.embed_dims = EMBED


class _HeadBase(nn.Module):
    def __init__(self, num_feature_levels, encoder, positional_encoding, in_channels, channels, in_index=-1,
                 align_corners=False, dropout_ratio=0.1, norm_cfg=None, loss_decode=None, **kwargs):
        super().__init__()
        if num_feature_levels != 1:
            raise NotImplementedError("libddp_b200 implements num_feature_levels=1 (every shipped DDP config)")
        self.num_feature_levels = num_feature_levels
        self.encoder = _EncoderParams(**dict(encoder))
        pe = dict(positional_encoding)
        assert "num_feats" in pe
        assert pe["num_feats"] * 2 == self.encoder.embed_dims, \
            f"embed_dims should be exactly 2 times of num_feats. Found {self.encoder.embed_dims} and {pe['num_feats']}."
        if pe.get("type") != "SinePositionalEncoding" or not pe.get("normalize", False) or pe.get("offset", 0.) != -0.5:
            raise NotImplementedError("libddp_b200 implements SinePositionalEncoding(normalize=True, offset=-0.5)")
        self.positional_encoding_cfg = pe
        self.embed_dims = self.encoder.embed_dims
        self.in_channels = list(in_channels) if isinstance(in_channels, (list, tuple)) else in_channels
        self.channels = channels
        self.in_index = in_index
        self.align_corners = align_corners
        self.dropout_ratio = dropout_ratio
        self.norm_cfg = norm_cfg
        self.loss_decode_cfg = loss_decode
        self.fp16_enabled = False
        self.input_transform = "multiple_select"
        self.__dict__['_owner'] = None          # weakref to the segmentor / depther that owns the engine

    def _init_transformer_weights(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, _MSDeformAttnParams):
                m.init_weights()

    def forward(self, inputs, times):
        """One denoiser evaluation (deformable_head_with_time.py:90-132): inputs [feat (rows,256,h,w)], times (b,1024)."""
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("the head is driven through its DDP segmentor/depther (which owns the CUDA engine)")
        return owner._head_forward(inputs[-1], times)

    def forward_test(self, inputs, times, img_metas, test_cfg):
        return self.forward(inputs, times)

    def forward_train(self, *a, **k):
        raise NotImplementedError("training is outside the scope of ddp_b200 (inference hot path only)")


@HEADS.register_module()
class DeformableHeadWithTime(_HeadBase):
    """Segmentation head: BaseDecodeHead surface (decode_head.py:58-140) + conv_seg (1x1)."""

    def __init__(self, num_feature_levels, encoder, positional_encoding, num_classes=None, out_channels=None,
                 threshold=None, ignore_index=255, **kwargs):
        super().__init__(num_feature_levels, encoder, positional_encoding, **kwargs)
        if num_classes is None:
            raise TypeError("DeformableHeadWithTime: missing num_classes")
        if out_channels is None:
            out_channels = num_classes
        if out_channels != num_classes:
            raise NotImplementedError("out_channels != num_classes (binary 1-channel output) is not built")
        self.num_classes = num_classes
        self.out_channels = out_channels
        self.threshold = threshold
        self.ignore_index = ignore_index
        self.conv_seg = nn.Conv2d(self.channels, self.out_channels, kernel_size=1)
        nn.init.normal_(self.conv_seg.weight, mean=0, std=0.01)
        nn.init.constant_(self.conv_seg.bias, 0)
        self._init_transformer_weights()


class DepthDeformableHeadWithTime(_HeadBase):
    """Depth head: DepthBaseDecodeHead surface (depth/.../decode_head.py:60-112) + conv_depth (3x3)."""

    def __init__(self, num_feature_levels, encoder, positional_encoding, min_depth=1e-3, max_depth=None, classify=False,
                 n_bins=256, scale_up=False, use_eps=True, init_inputs=False, **kwargs):
        super().__init__(num_feature_levels, encoder, positional_encoding, **kwargs)
        if classify or scale_up or not use_eps:
            raise NotImplementedError("libddp_b200 implements depth_pred = relu(conv3x3) + min_depth "
                                      "(classify=False, scale_up=False, use_eps=True: every shipped config)")
        if max_depth is None:
            raise TypeError("DeformableHeadWithTime: missing max_depth")
        self.min_depth, self.max_depth = min_depth, max_depth
        self.classify, self.n_bins, self.scale_up, self.use_eps = classify, n_bins, scale_up, use_eps
        self.num_classes = 1
        self.conv_depth = nn.Conv2d(self.channels, 1, kernel_size=3, padding=1, stride=1)
        self._init_transformer_weights()
