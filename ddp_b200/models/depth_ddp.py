"""Depth ``DDP`` depther — plug-in surface of depth/depth/models/depther/ddp.py:34-250."""
import weakref

import torch
import torch.nn as nn

from ..registry import DEPTHER
from .ddp import _DiffusionSegmentorBase, _ConvModule1x1, LearnedSinusoidalPosEmb, resize, EMBED
from .deformable_head_with_time import DepthDeformableHeadWithTime


@DEPTHER.register_module(name="DDP")
class DDP(_DiffusionSegmentorBase):
    def __init__(self, bit_scale=1, bits=8, timesteps=1, randsteps=1, time_difference=1, learned_sinusoidal_dim=16,
                 sample_range=(0, 0.999), ddim=True, rule=None, min_depth=1e-3, max_depth=80, backbone=None,
                 decode_head=None, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None, pretrained=None,
                 init_cfg=None, gemm_mode="tc_3xf16"):
        super().__init__()
        self._engine = None
        self.gemm_mode = gemm_mode
        self._init_encoder(backbone, neck, pretrained)
        head_cfg = dict(decode_head)
        if head_cfg.pop("type") != "DeformableHeadWithTime":
            raise NotImplementedError("decode_head.type must be DeformableHeadWithTime")
        for k in ("loss_decode",):
            head_cfg.setdefault(k, None)
        self.decode_head = DepthDeformableHeadWithTime(**head_cfg)
        self.decode_head.__dict__['_owner'] = weakref.ref(self)     # not a submodule: no cycle in state_dict
        self.align_corners = self.decode_head.align_corners
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.bit_scale, self.BITS = bit_scale, bits
        self.timesteps, self.randsteps = timesteps, randsteps
        self.time_difference, self.sample_range = time_difference, sample_range
        if not ddim:
            raise NotImplementedError("ddpm_step is not defined by the reference either (depth/.../ddp.py:244)")
        self.ddim = ddim
        self.min_depth, self.max_depth = min_depth, max_depth
        self.learned_sinusoidal_dim = learned_sinusoidal_dim
        print("sample range:", sample_range)
        print("timesteps: {}, randsteps: {}".format(timesteps, randsteps))
        c = self.decode_head.in_channels[0]
        if c != EMBED:
            raise NotImplementedError("libddp_b200 is built for 256-channel neck features")
        self.down = _ConvModule1x1(c + 1, c)
        time_dim = c * 4
        self.time_mlp = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinusoidal_dim),
                                      nn.Linear(learned_sinusoidal_dim + 1, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))

    def _engine_kwargs(self):
        # the clamp of encode_decode uses the HEAD's min/max depth, the normalisation the depther's (ddp.py:103, 240)
        if (self.decode_head.min_depth, self.decode_head.max_depth) != (self.min_depth, self.max_depth):
            raise NotImplementedError("decode_head and depther min/max depth differ")
        return dict(task="depth", timesteps=self.timesteps, time_difference=self.time_difference,
                    sample_range=self.sample_range, bit_scale=self.bit_scale,
                    learned_sinusoidal_dim=self.learned_sinusoidal_dim, num_layers=self.decode_head.encoder.num_layers,
                    min_depth=self.min_depth, max_depth=self.max_depth)

    def encode_decode(self, img, img_metas, rescale=False):
        """depth/.../ddp.py:97-110 (the clamp runs inside the library's last step)."""
        x = self.extract_feat(img)[0]
        out = self.sample(x, img_metas, _clamped=True)
        if rescale:
            out = resize(input=out, size=img.shape[2:], mode="bilinear", align_corners=self.align_corners)
        return out

    @torch.no_grad()
    def sample(self, x, img_metas=None, noise=None, _clamped=False):
        """depth/.../ddp.py:229-247, batched; returns the mean prediction clamped to [min_depth, max_depth]
        (the reference applies that clamp one line later, in encode_decode)."""
        b, c, h, w = x.shape
        if noise is None:
            noise = torch.randn((b, self.randsteps, 1, h, w), device=x.device)
        return self.engine().sample(x.float(), noise)

    def _head_forward(self, feat, times):
        """decode_head.forward(inputs, times): one denoiser call through ddp_head_forward."""
        if times.shape[0] != 1 and not bool((times == times[0:1]).all()):
            raise NotImplementedError("one time embedding per call (the sampling loop uses the same t for every row)")
        return self.engine().head_forward(feat.float(), times[0])

    def simple_test(self, img, img_meta, rescale=True):
        depth_pred = self.encode_decode(img, img_meta, rescale)
        return list(depth_pred.cpu().numpy())
