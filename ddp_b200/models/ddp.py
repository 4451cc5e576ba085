"""``DDP`` / ``SelfAlignedDDP`` segmentors — plug-in surface of the reference, hot path in libddp_b200.so.

Reference: segmentation/mmseg/models/segmentors/ddp.py:49-290, self_aligned_ddp.py (inference identical),
encoder_decoder.py:24-304, base.py:62-110.  Same constructor signature, methods and state-dict keys; the
T-step loop (`ddim_sample`) is ONE call into the CUDA library instead of ~180 PyTorch launches per step.
"""
import warnings

import weakref

import torch
import torch.nn as nn
import torch.nn.functional as F

from .._stale import fingerprint
from ..engine import DecodeEngine
from ..registry import SEGMENTORS, MODELS, build_head

EMBED = 256
_COLD_PREFIXES = ("backbone.", "neck.", "auxiliary_head.")      # not weights of the decode loop


def resize(input, size=None, scale_factor=None, mode="nearest", align_corners=None, warning=True):
    """segmentation/mmseg/ops/wrappers.py:8-27."""
    return F.interpolate(input, size, scale_factor, mode, align_corners)


class LearnedSinusoidalPosEmb(nn.Module):
    """Parameter container of ddp.py:31-46 (key ``time_mlp.0.weights``)."""

    def __init__(self, dim):
        super().__init__()
        assert (dim % 2) == 0
        self.weights = nn.Parameter(torch.randn(dim // 2))


class _ConvModule1x1(nn.Module):
    """mmcv ConvModule(k=1, no norm, no act): key ``<name>.conv.{weight,bias}`` (ddp.py:92-100)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, padding=0)


class _MissingEncoder(nn.Module):
    """Placeholder for a backbone / neck type that is not registered here (the encoder is outside the hot path)."""

    def __init__(self, type, **kw):
        super().__init__()
        self.missing_type = type

    def forward(self, *a, **k):
        raise RuntimeError(f"encoder module '{self.missing_type}' is not available in ddp_b200 (out of scope): "
                           "register it in ddp_b200.registry.MODELS, run inside mmseg, or pass neck features to ddim_sample")


def _build_encoder_part(cfg):
    if cfg is None:
        return None
    if isinstance(cfg, (list, tuple)):
        from ..neck import fuse_neck       # neck=[FPN, MultiStageMerging] runs as one library call
        return fuse_neck([_build_encoder_part(c) for c in cfg])
    typ = cfg.get("type")
    if isinstance(typ, str) and typ not in MODELS:
        from ..registry import host_framework_builder
        host = host_framework_builder(typ)          # inside real mmseg / depth: the reference's own backbone / neck classes
        if host is not None:
            return host.build(cfg)
        warnings.warn(f"'{typ}' is not registered; using a placeholder (the encoder is outside the ddp_b200 hot path)")
        return _MissingEncoder(**cfg)
    return MODELS.build(cfg)


class _DiffusionSegmentorBase(nn.Module):
    """EncoderDecoder surface shared by the seg and depth plug-ins."""

    engine_task = "seg"

    def _init_encoder(self, backbone, neck, pretrained):
        if pretrained is not None:
            assert backbone.get("pretrained") is None, "both backbone and segmentor set pretrained weight"
        self.backbone = _build_encoder_part(backbone)
        if neck is not None:
            self.neck = _build_encoder_part(neck)

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    @property
    def with_decode_head(self):
        return hasattr(self, "decode_head") and self.decode_head is not None

    @property
    def with_auxiliary_head(self):          # training-only heads are not built (base.py:26-29 reads the attribute)
        return getattr(self, "auxiliary_head", None) is not None

    @staticmethod
    def right_pad_dims_to(x, t):
        """ddp.py:198-202: t with trailing singleton axes up to x's rank."""
        extra = x.ndim - t.ndim
        return t if extra <= 0 else t.reshape(*t.shape, *([1] * extra))

    def _sampling_time_pairs(self):
        from .. import schedule as S
        return S.sampling_timesteps_seg(self.timesteps, self.time_difference, self.sample_range)

    def _get_sampling_timesteps(self, batch, *, device):
        """ddp.py:204-213 / depther/ddp.py:210-218: the (t_now, t_next) pairs of the T steps as (2, batch) tensors — the same
        host function (`ddp_b200.schedule`) the library's per-step constants are computed from."""
        return [torch.tensor(pair, device=device)[:, None].repeat(1, batch) for pair in self._sampling_time_pairs()]

    def extract_feat(self, img):
        x = self.backbone(img)
        if self.with_neck:
            x = self.neck(x)
        return x

    # ---- engine management -----------------------------------------------------------------
    def _hot_state_dict(self):
        return {k: v for k, v in self.state_dict().items() if not k.startswith(_COLD_PREFIXES)}

    def _engine_kwargs(self):
        raise NotImplementedError

    def engine(self) -> DecodeEngine:
        """The CUDA decode engine bound to the current parameters (rebuilt after load_state_dict / refresh)."""
        fp = fingerprint(self, _COLD_PREFIXES)
        if self._engine is None or fp != getattr(self, "_engine_fp", None):
            eng = DecodeEngine(gemm_mode=self.gemm_mode, **self._engine_kwargs())
            eng.load_state_dict(self._hot_state_dict())
            self._engine, self._engine_fp = eng, fp
        return self._engine

    def refresh_engine(self):
        self._engine = None

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.refresh_engine()
        return out

    def _apply(self, fn, *a, **k):          # .cuda() / .to(): parameters moved, engine re-reads them lazily
        self.refresh_engine()
        return super()._apply(fn, *a, **k)

    def forward_train(self, *a, **k):
        raise NotImplementedError("training is outside the scope of ddp_b200 (inference hot path only)")

    def forward_dummy(self, img):
        """encoder_decoder.py:142-146 (tools/get_flops.py, tools/benchmark.py style callers): one encode_decode."""
        return self.encode_decode(img, None)

    def forward(self, img, img_metas, return_loss=False, **kwargs):
        """base.py:96-110."""
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:62-94 (single-augmentation path; aug_test averages `inference` outputs)."""
        for var, name in [(imgs, "imgs"), (img_metas, "img_metas")]:
            if not isinstance(var, list):
                raise TypeError(f"{name} must be a list, but got {type(var)}")
        if len(imgs) != len(img_metas):
            raise ValueError(f"num of augmentations ({len(imgs)}) != num of image meta ({len(img_metas)})")
        if len(imgs) == 1:
            return self.simple_test(imgs[0], img_metas[0], **kwargs)
        return self.aug_test(imgs, img_metas, **kwargs)


@SEGMENTORS.register_module()
class DDP(_DiffusionSegmentorBase):
    def __init__(self, bit_scale=0.1, timesteps=1, randsteps=1, time_difference=1, learned_sinusoidal_dim=16,
                 sample_range=(0, 0.999), noise_schedule="cosine", diffusion="ddim", accumulation=False,
                 backbone=None, decode_head=None, neck=None, auxiliary_head=None, train_cfg=None, test_cfg=None,
                 pretrained=None, init_cfg=None, gemm_mode="tc_3xf16"):
        super().__init__()
        self._engine = None
        self.gemm_mode = gemm_mode
        self.fused_tail = True          # simple_test: fused resize+softmax+argmax kernel when the shapes allow it
        self._init_encoder(backbone, neck, pretrained)
        self.decode_head = build_head(decode_head)
        self.decode_head.__dict__['_owner'] = weakref.ref(self)     # not a submodule: no cycle in state_dict
        self.align_corners = self.decode_head.align_corners
        self.num_classes = self.decode_head.num_classes
        self.out_channels = self.decode_head.out_channels
        # the auxiliary head only matters for training (deep supervision); its config is accepted and ignored
        self.auxiliary_head_cfg = auxiliary_head
        self.train_cfg, self.test_cfg = train_cfg, test_cfg

        self.bit_scale = bit_scale
        self.timesteps = timesteps
        self.randsteps = randsteps
        self.diffusion = diffusion
        self.time_difference = time_difference
        self.sample_range = sample_range
        self.use_gt = False
        self.accumulation = accumulation
        self.learned_sinusoidal_dim = learned_sinusoidal_dim
        self.embedding_table = nn.Embedding(self.num_classes + 1, self.decode_head.in_channels[0])
        print(f" timesteps: {timesteps}, randsteps: {randsteps}, sample_range: {sample_range}, diffusion: {diffusion}")
        if noise_schedule not in ("linear", "cosine"):
            raise ValueError(f"invalid noise schedule {noise_schedule}")
        if diffusion not in ("ddim", "ddpm"):
            pass        # like the reference, an unknown sampler only fails when encode_decode is called (ddp.py:123)
        self.noise_schedule = noise_schedule
        c = self.decode_head.in_channels[0]
        if c != EMBED:
            raise NotImplementedError("libddp_b200 is built for 256-channel neck features")
        self.transform = _ConvModule1x1(c * 2, c)
        time_dim = c * 4
        self.time_mlp = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinusoidal_dim),
                                      nn.Linear(learned_sinusoidal_dim + 1, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))

    def _engine_kwargs(self):
        return dict(task="seg", num_classes=self.num_classes, timesteps=self.timesteps,
                    time_difference=self.time_difference, sample_range=self.sample_range,
                    noise_schedule=self.noise_schedule, diffusion=self.diffusion, accumulation=self.accumulation,
                    bit_scale=self.bit_scale, learned_sinusoidal_dim=self.learned_sinusoidal_dim,
                    num_layers=self.decode_head.encoder.num_layers)

    # ---- hot path ---------------------------------------------------------------------------------
    def encode_decode(self, img, img_metas):
        """ddp.py:114-129."""
        x = self.extract_feat(img)[0]
        if self.diffusion == "ddim":
            out = self.ddim_sample(x, img_metas)
        elif self.diffusion == "ddpm":
            out = self.ddpm_sample(x, img_metas)
        else:
            raise NotImplementedError
        return resize(input=out, size=img.shape[2:], mode="bilinear", align_corners=self.align_corners)

    @torch.no_grad()
    def ddim_sample(self, x, img_metas=None, noise=None):
        """ddp.py:215-246, batched: x (b,256,h,w) -> (b,C,h,w).  The reference loop is only defined for b=1
        (its batch axis is ``randsteps``); here every image gets its own ``randsteps`` samples and its own mean.
        ``noise`` (b, randsteps, 256, h, w) may be given; otherwise it is drawn like ddp.py:220 does."""
        b, c, h, w = x.shape
        if noise is None:
            noise = torch.randn((b, self.randsteps, c, h, w), device=x.device)
        return self.engine().sample(x.float(), noise)

    @torch.no_grad()
    def ddpm_sample(self, x, img_metas=None, noise=None, step_noise=None):
        """ddp.py:248-290, batched.  The noise is drawn in the reference's order: the initial mask_t, then
        randn_like(mask_t) at every step."""
        b, c, h, w = x.shape
        if noise is None:
            noise = torch.randn((b, self.randsteps, c, h, w), device=x.device)
        if step_noise is None:
            step_noise = torch.stack([torch.randn_like(noise) for _ in range(self.timesteps)])
        return self.engine().sample(x.float(), noise, step_noise=step_noise)

    def _head_forward(self, feat, times):
        """decode_head.forward(inputs, times): one denoiser call through ddp_head_forward."""
        if times.shape[0] != 1 and not bool((times == times[0:1]).all()):
            raise NotImplementedError("one time embedding per call (the sampling loop uses the same t for every row)")
        return self.engine().head_forward(feat.float(), times[0])

    def _decode_head_forward_test(self, x, t, img_metas):
        return self.decode_head.forward_test(x, t, img_metas, self.test_cfg)

    # ---- callers of the hot path (encoder_decoder.py:229-304) -------------------------------------
    def whole_inference(self, img, img_meta, rescale):
        seg_logit = self.encode_decode(img, img_meta)
        if rescale:
            resize_shape = img_meta[0]["img_shape"][:2]
            seg_logit = seg_logit[:, :, :resize_shape[0], :resize_shape[1]]
            size = img_meta[0]["ori_shape"][:2]
            seg_logit = resize(seg_logit, size=size, mode="bilinear", align_corners=self.align_corners, warning=False)
        return seg_logit

    def slide_inference(self, img, img_meta, rescale):
        """encoder_decoder.py:181-227: overlapping crop_size windows at `stride`, each decoded by encode_decode, averaged
        where they overlap (no shipped DDP config selects it; EncoderDecoder callers may).  Every window has the same size
        (min(crop, image)) and windows are independent, so `test_cfg.window_batch` of them (default 8) are stacked along the
        batch axis and decoded by ONE encode_decode call instead of one call per window; the results are added in the
        reference's (row, column) window order, so equal window logits give bit-equal sums."""
        cfg = self.test_cfg

        def pair(v):
            return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))

        h_stride, w_stride = pair(cfg["stride"])
        h_crop, w_crop = pair(cfg["crop_size"])
        per_call = max(1, int(cfg.get("window_batch", 8)))
        batch, _, h_img, w_img = img.shape
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        boxes = []
        for hi in range(h_grids):
            for wi in range(w_grids):
                y2 = min(hi * h_stride + h_crop, h_img)
                x2 = min(wi * w_stride + w_crop, w_img)
                boxes.append((max(y2 - h_crop, 0), y2, max(x2 - w_crop, 0), x2))
        preds = img.new_zeros((batch, self.out_channels, h_img, w_img))
        count = img.new_zeros((batch, 1, h_img, w_img))
        for i in range(0, len(boxes), per_call):
            group = boxes[i:i + per_call]
            crops = torch.cat([img[:, :, y1:y2, x1:x2] for y1, y2, x1, x2 in group], dim=0)
            logits = self.encode_decode(crops, img_meta)
            for k, (y1, y2, x1, x2) in enumerate(group):
                preds[:, :, y1:y2, x1:x2] += logits[k * batch:(k + 1) * batch]
                count[:, :, y1:y2, x1:x2] += 1
        assert int((count == 0).sum()) == 0
        preds = preds / count
        if rescale:
            resize_shape = img_meta[0]["img_shape"][:2]
            preds = preds[:, :, :resize_shape[0], :resize_shape[1]]
            preds = resize(preds, size=img_meta[0]["ori_shape"][:2], mode="bilinear", align_corners=self.align_corners,
                           warning=False)
        return preds

    def _fused_view(self, img, img_meta, rescale, accum=None):
        """encode_decode + whole_inference + softmax + flip of `inference` for one view as ONE kernel after the loop
        (ddp_tail_probs): probabilities (b,C,H,W), added into `accum` when given.  None when the view needs the eager path
        (align_corners=True heads, the ddpm sampler's own noise bookkeeping, a head with one output channel)."""
        if not self.fused_tail or self.align_corners or self.out_channels == 1 or (self.test_cfg or {}).get("mode", "whole") != "whole":
            return None
        x = self.extract_feat(img)[0]
        logits = self.ddim_sample(x, img_meta) if self.diffusion == "ddim" else self.ddpm_sample(x, img_meta)
        meta = img_meta[0]
        flip = meta.get("flip_direction", "horizontal") if meta.get("flip", False) else None
        assert flip in (None, "horizontal", "vertical")
        if rescale:
            return self.engine().tail_probs(logits, img.shape[2:], crop=meta["img_shape"][:2], out_size=meta["ori_shape"][:2],
                                            flip=flip, accum=accum)
        return self.engine().tail_probs(logits, img.shape[2:], flip=flip, accum=accum)

    def inference(self, img, img_meta, rescale):
        mode = (self.test_cfg or {}).get("mode", "whole")
        assert mode in ["slide", "whole"]
        ori_shape = img_meta[0]["ori_shape"] if "ori_shape" in img_meta[0] else None
        assert all(m.get("ori_shape") == ori_shape for m in img_meta)
        fused = self._fused_view(img, img_meta, rescale)      # None in slide mode
        if fused is not None:
            return fused
        if mode == "slide":
            seg_logit = self.slide_inference(img, img_meta, rescale)
        else:
            seg_logit = self.whole_inference(img, img_meta, rescale)
        output = F.softmax(seg_logit, dim=1)
        if img_meta[0].get("flip", False):
            flip_direction = img_meta[0]["flip_direction"]
            assert flip_direction in ["horizontal", "vertical"]
            output = output.flip(dims=(3,)) if flip_direction == "horizontal" else output.flip(dims=(2,))
        return output

    def simple_test(self, img, img_meta, rescale=True):
        meta = img_meta[0]
        # k_resize_argmax implements the align_corners=False source-index mapping (every DDP config); a head configured
        # with align_corners=True takes the unfused path below, which honours the flag (encoder_decoder.py:241-249)
        plain = (not self.align_corners and not meta.get("flip", False) and (self.test_cfg or {}).get("mode", "whole") == "whole"
                 and tuple(meta.get("img_shape", ())[:2]) == tuple(img.shape[2:])
                 and (not rescale or tuple(meta.get("ori_shape", ())[:2]) == tuple(img.shape[2:])))
        if self.fused_tail and plain and self.diffusion == "ddim":
            # resize + softmax + argmax of encode_decode / whole_inference / inference in ONE kernel (the second
            # resize of whole_inference is the identity when ori_shape == img_shape)
            x = self.extract_feat(img)[0]
            logits = self.ddim_sample(x, img_meta)
            return list(self.engine().resize_argmax(logits, img.shape[2:]).cpu().numpy().astype("int64"))
        seg_logit = self.inference(img, img_meta, rescale)
        seg_pred = seg_logit.argmax(dim=1)
        return list(seg_pred.cpu().numpy())

    def aug_test(self, imgs, img_metas, rescale=True):
        """encoder_decoder.py:295-304: the views' probability maps are summed (here: accumulated in place by the tail kernel,
        one launch per view), divided by their number and arg-maxed (the division does not change the argmax)."""
        assert rescale
        seg_logit = self.inference(imgs[0], img_metas[0], rescale)
        fused = self.fused_tail and not self.align_corners and self.out_channels != 1
        for i in range(1, len(imgs)):
            if fused and self._fused_view(imgs[i], img_metas[i], rescale, accum=seg_logit) is not None:
                continue
            seg_logit += self.inference(imgs[i], img_metas[i], rescale)
        if fused:
            return list(self.engine().probs_argmax(seg_logit).cpu().numpy().astype("int64"))
        seg_logit /= len(imgs)
        return list(seg_logit.argmax(dim=1).cpu().numpy())


@SEGMENTORS.register_module()
class SelfAlignedDDP(DDP):
    """self_aligned_ddp.py: differs from DDP only in forward_train (lines 148-175); inference is the same kernel."""
