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

    def _sampling_time_pairs(self):
        from .. import schedule as S
        return S.sampling_timesteps_depth(self.timesteps, self.time_difference)

    @staticmethod
    def gamma(t, ns=0.0002, ds=0.00025):
        """depther/ddp.py:207-208 (the library receives sqrt(gamma) per step from the same host function)."""
        from .. import schedule as S
        return S.gamma(t, ns, ds)

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

    # ---- callers of the hot path (depth/depth/models/depther/encoder_decoder.py:156-229) ------------
    def whole_inference(self, img, img_meta, rescale):
        return self.encode_decode(img, img_meta, rescale)

    @staticmethod
    def _unflip(pred, meta):
        """encoder_decoder.py:187-195: a flipped view's prediction is flipped back."""
        if meta.get("flip", False):
            direction = meta["flip_direction"]
            assert direction in ["horizontal", "vertical"]
            pred = pred.flip(dims=(3,)) if direction == "horizontal" else pred.flip(dims=(2,))
        return pred

    def _check_view(self, img_meta):
        mode = (self.test_cfg or {}).get("mode", "whole")
        assert mode in ["slide", "whole"]
        ori_shape = img_meta[0]["ori_shape"]
        assert all(m["ori_shape"] == ori_shape for m in img_meta)
        if mode == "slide":
            raise NotImplementedError          # the reference raises here too (encoder_decoder.py:182-183)

    def inference(self, img, img_meta, rescale):
        """encoder_decoder.py:163-197."""
        self._check_view(img_meta)
        return self._unflip(self.whole_inference(img, img_meta, rescale), img_meta[0])

    def simple_test(self, img, img_meta, rescale=True):
        """encoder_decoder.py:198-208."""
        return list(self.inference(img, img_meta, rescale).cpu().numpy())

    def aug_test(self, imgs, img_metas, rescale=True):
        """encoder_decoder.py:210-229: mean of the views' (un-flipped) predictions.  The shipped NYU / KITTI test pipelines
        are MultiScaleFlipAug with flip=True, i.e. TWO views of the same size per image: consecutive views of equal shape
        are stacked along the batch axis and decoded by ONE encode_decode call (`test_cfg.batch_views`, default True) —
        views are independent, and two images per call fill the GPU better than two calls of one.  Summed in view order."""
        assert rescale
        n, b = len(imgs), imgs[0].shape[0]
        batch_views = bool((self.test_cfg or {}).get("batch_views", True))
        preds, i = [None] * n, 0
        while i < n:
            j = i + 1
            while batch_views and j < n and imgs[j].shape == imgs[i].shape:
                j += 1
            if j - i == 1:
                preds[i] = self.inference(imgs[i], img_metas[i], rescale)
            else:
                for k in range(i, j):
                    self._check_view(img_metas[k])
                both = self.whole_inference(torch.cat(imgs[i:j], dim=0), img_metas[i], rescale)
                for k in range(i, j):
                    preds[k] = self._unflip(both[(k - i) * b:(k - i + 1) * b], img_metas[k][0])
            i = j
        depth_pred = preds[0].clone()
        for k in range(1, n):
            depth_pred += preds[k]
        depth_pred /= n
        return list(depth_pred.cpu().numpy())
