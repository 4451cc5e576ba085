"""BEV map segmentation: the BEV tree's ``DDP`` fusion model and its ``DeformableHeadWithTime`` over the C ABI.

Reference: bev/mmdet3d/models/fusion_models/ddp.py:65-116 (constructor), :268-301 (``ddim_sample``);
bev/mmdet3d/models/heads/segm/deformable_head_with_time.py:58-98 (``BEVGridTransform``), :100-142 (head constructor),
:178-241 (head forward).  Same constructor arguments and state-dict keys (``embedding_table.weight``,
``transform.conv.*``, ``time_mlp.*``, ``heads.map.encoder.layers.N.*``, ``heads.map.conv_seg.*``); the sampling loop is ONE
call into libddp_b200.so (``ddp_bev_sample``), which runs the same denoiser kernels as the segmentation path on the
output map grid.  The sensor encoders, the fuser and the BEV decoder of ``BEVFusion`` are outside the hot path and are
not built: ``ddim_sample`` takes the fused BEV feature ``x`` exactly as the reference's does.  No tensor math happens in
Python; there is no CPU fallback.
"""
import ctypes
from typing import Mapping

import torch
import torch.nn as nn

from . import _lib as L
from . import schedule as S
from ._stale import fingerprint
from .models.ddp import LearnedSinusoidalPosEmb, _ConvModule1x1
from .models.deformable_head_with_time import _EncoderParams, _MSDeformAttnParams
from .registry import Registry

FUSIONMODELS = Registry("fusion_models")      # bev/mmdet3d/models/builder.py keeps separate registries
BEV_HEADS = Registry("bev_heads")
NUM_CLASSES = 6                               # fusion_models/ddp.py:89


def _fptr(seq):
    return (ctypes.c_float * len(seq))(*seq)


def grid_coords(input_scope, output_scope):
    """Normalised sampling coordinates of BEVGridTransform, with the reference's own torch ops
    (heads/segm/deformable_head_with_time.py:82-88) so that they are bit-identical to a reference run."""
    coords = []
    for (imin, imax, _), (omin, omax, ostep) in zip(input_scope, output_scope):
        v = torch.arange(omin + ostep / 2, omax, ostep)
        coords.append(((v - imin) / (imax - imin) * 2 - 1).to(torch.float32))
    return coords


class BevDecodeEngine:
    """Python handle of one ``ddp_bev`` bound to one CUDA device."""

    def __init__(self, timesteps=3, time_difference=1, noise_schedule="cosine", diffusion="ddim", bit_scale=0.01,
                 threshold=0.5, feat_channels=512, learned_sinusoidal_dim=16, num_layers=5, sample_range=(0, 0.999),
                 gemm_mode="tc_3xf16", device=None, host_schedule=True):
        if noise_schedule not in ("cosine", "linear"):
            raise ValueError(f"invalid noise schedule {noise_schedule}")               # fusion_models/ddp.py:102
        if diffusion != "ddim":
            raise NotImplementedError(f"unsupported diffusion: {diffusion} (the BEV reference's ddpm_sample cannot run)")
        if gemm_mode not in L.GEMM_MODES:
            raise ValueError(f"gemm_mode must be one of {sorted(L.GEMM_MODES)}")
        if not torch.cuda.is_available():
            raise RuntimeError("ddp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self.timesteps, self.time_difference, self.noise_schedule = timesteps, time_difference, noise_schedule
        self.feat_channels = feat_channels
        cfg = L.DDPBevConfig(abi_version=L.ABI_VERSION, timesteps=timesteps, time_difference=time_difference,
                             noise_schedule=L.SCHEDULE_COSINE if noise_schedule == "cosine" else L.SCHEDULE_LINEAR,
                             diffusion=L.DIFFUSION_DDIM, learned_sinusoidal_dim=learned_sinusoidal_dim,
                             num_layers=num_layers, feat_channels=feat_channels, gemm_mode=L.GEMM_MODES[gemm_mode],
                             bit_scale=float(bit_scale), threshold=float(threshold))
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.ddp_bev_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_bev_last_error(None).decode())
        self._plan = None
        self._ws = None
        if host_schedule:
            # the BEV time pairs are the segmentation pairs with sample_range[0] = 0 (fusion_models/ddp.py:128-136)
            l, a, s, an, sn = S.seg_schedule(timesteps, time_difference, (0, sample_range[1]), noise_schedule)
            self._check(self.lib.ddp_bev_set_schedule(self._h, timesteps, _fptr(l), _fptr(a), _fptr(s), _fptr(an), _fptr(sn)))

    def _check(self, rc):
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_bev_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.ddp_bev_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_names(self):
        out, n = {}, ctypes.c_int64()
        for i in range(self.lib.ddp_bev_weight_count(self._h)):
            # two statements on purpose: `out[f(byref(n))] = n.value` reads n.value BEFORE the call (Python evaluates the
            # right-hand side first) — the round-1 bug that failed every neck / BEV hardware test
            name = self.lib.ddp_bev_weight_name(self._h, i, ctypes.byref(n)).decode()
            out[name] = n.value
        return out

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]):
        need = self.weight_names()
        missing = [k for k in need if k not in sd]
        if missing:
            raise KeyError(f"state dict lacks BEV decode-path weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, numel in need.items():
            t = sd[k].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != numel:
                raise ValueError(f"{k}: expected {numel} elements, got {tuple(t.shape)}")
            self._check(self.lib.ddp_bev_set_weight(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), numel))
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_bev_commit_weights(self._h))
        self._plan = None

    def plan(self, B, R, h, w, grid_y, grid_x):
        gy = grid_y.detach().to("cpu", torch.float32).contiguous()
        gx = grid_x.detach().to("cpu", torch.float32).contiguous()
        key = (B, R, h, w, tuple(gy.tolist()), tuple(gx.tolist()))
        if self._plan == key:
            return
        nbytes = ctypes.c_size_t()
        fp = ctypes.POINTER(ctypes.c_float)
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_bev_plan(self._h, B, R, h, w, gy.numel(), gx.numel(),
                                              ctypes.cast(gy.data_ptr(), fp), ctypes.cast(gx.data_ptr(), fp),
                                              ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value + 256:
            self._ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
        self._ws_bytes = nbytes.value
        self._plan = key
        self._out_hw = (gy.numel(), gx.numel())

    def _ws_ptr(self):
        return (self._ws.data_ptr() + 255) // 256 * 256          # the library wants a 256-byte aligned workspace

    @property
    def last_launch_count(self):
        return int(self.lib.ddp_bev_last_launch_count(self._h))

    def sample(self, x, noise, grid_y, grid_x):
        """x (B,feat,h,w), noise (B,R,256,h,w) CUDA fp32 -> (B,6,H',W') mean sigmoid maps."""
        if x.dim() != 4 or x.shape[1] != self.feat_channels:
            raise ValueError(f"x has shape {tuple(x.shape)}, expected (B, {self.feat_channels}, h, w)")
        B, _, h, w = x.shape
        if noise.dim() != 5 or noise.shape[0] != B or noise.shape[2] != 256 or tuple(noise.shape[3:]) != (h, w):
            raise ValueError(f"noise has shape {tuple(noise.shape)}, expected ({B}, R, 256, {h}, {w})")
        if x.device != self.device or noise.device != self.device:
            raise ValueError(f"inputs must be on {self.device}")
        R = noise.shape[1]
        if B == 0:
            return x.new_empty((0, NUM_CLASSES, grid_y.numel(), grid_x.numel()))
        self.plan(B, R, h, w, grid_y, grid_x)
        x = x.detach().to(torch.float32).contiguous()
        noise = noise.detach().to(torch.float32).contiguous()
        out = torch.empty((B, NUM_CLASSES) + self._out_hw, dtype=torch.float32, device=self.device)
        vp = ctypes.c_void_p
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_bev_sample(self._h, vp(x.data_ptr()), vp(noise.data_ptr()), vp(out.data_ptr()),
                                                vp(self._ws_ptr()), self._ws_bytes, vp(stream)))
        return out


@BEV_HEADS.register_module(name="DeformableHeadWithTime")
class BevDeformableHeadWithTime(nn.Module):
    """Parameter container + geometry of the BEV head (heads/segm/deformable_head_with_time.py:100-142)."""

    def __init__(self, num_feature_levels, encoder, positional_encoding, classes, loss, grid_transform, in_channels=256,
                 seg_conv_kernel=1):
        super().__init__()
        if num_feature_levels != 1:
            raise NotImplementedError("libddp_b200 implements num_feature_levels=1")
        if seg_conv_kernel != 1:
            raise NotImplementedError("libddp_b200 implements the 1x1 conv_seg of the shipped BEV configs")
        if len(classes) != NUM_CLASSES:
            raise NotImplementedError(f"the BEV DDP model hard-codes {NUM_CLASSES} map classes (fusion_models/ddp.py:89)")
        if grid_transform.get("prescale_factor", 1) != 1:
            raise NotImplementedError("BEVGridTransform(prescale_factor != 1) is not built")
        self.num_feature_levels = num_feature_levels
        self.encoder = _EncoderParams(**dict(encoder))
        pe = dict(positional_encoding)
        assert "num_feats" in pe
        assert pe["num_feats"] * 2 == self.encoder.embed_dims, \
            f"embed_dims should be exactly 2 times of num_feats. Found {self.encoder.embed_dims} and {pe['num_feats']}."
        self.embed_dims = self.encoder.embed_dims
        self.classes = list(classes)
        self.loss = loss
        self.input_scope = [tuple(s) for s in grid_transform["input_scope"]]
        self.output_scope = [tuple(s) for s in grid_transform["output_scope"]]
        self.conv_seg = nn.Conv2d(in_channels, len(classes), kernel_size=1)
        for p in self.parameters():                                    # init_weights(), :144-151
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, _MSDeformAttnParams):
                m.init_weights()

    def grid_coords(self):
        return grid_coords(self.input_scope, self.output_scope)

    def forward(self, inputs, times, target=None):
        raise NotImplementedError("the BEV head is driven through DDP.ddim_sample (one library call for the whole loop)")


@FUSIONMODELS.register_module(name="DDP")
class BevDDP(nn.Module):
    """bev/mmdet3d/models/fusion_models/ddp.py:65-116, :268-301.  ``encoders`` / ``fuser`` / ``decoder`` (the BEVFusion
    encoder side) are accepted and ignored: they are outside the hot path."""

    def __init__(self, bit_scale=1, timesteps=1, randsteps=1, time_difference=1, learned_sinusoidal_dim=16,
                 sample_range=(0, 0.999), noise_schedule="cosine", diffusion="ddim", threshold=0.5, feat_channels=512,
                 tmp_channels=256, gemm_mode="tc_3xf16", heads=None, **kwargs):
        super().__init__()
        if noise_schedule not in ("linear", "cosine"):
            raise ValueError(f"invalid noise schedule {noise_schedule}")               # :102
        if tmp_channels != 256:
            raise NotImplementedError("libddp_b200 is built for tmp_channels=256")
        self.bit_scale, self.timesteps, self.randsteps = bit_scale, timesteps, randsteps
        self.diffusion, self.time_difference, self.sample_range = diffusion, time_difference, sample_range
        self.noise_schedule, self.threshold, self.feat_channels = noise_schedule, threshold, feat_channels
        self.learned_sinusoidal_dim = learned_sinusoidal_dim
        self.num_classes = NUM_CLASSES
        self.gemm_mode = gemm_mode
        self.embedding_table = nn.Embedding(self.num_classes + 1, tmp_channels)
        self.transform = _ConvModule1x1(tmp_channels + feat_channels, tmp_channels)
        time_dim = tmp_channels * 4
        self.time_mlp = nn.Sequential(LearnedSinusoidalPosEmb(learned_sinusoidal_dim),
                                      nn.Linear(learned_sinusoidal_dim + 1, time_dim), nn.GELU(), nn.Linear(time_dim, time_dim))
        self.heads = nn.ModuleDict()
        for name, cfg in (heads or {}).items():
            if cfg is not None:
                if name != "map":
                    raise NotImplementedError(f"head '{name}' is outside the scope of ddp_b200 (only the diffusion map head is built)")
                self.heads[name] = cfg if isinstance(cfg, nn.Module) else BEV_HEADS.build(dict(cfg))
        self._engine = None

    def engine(self) -> BevDecodeEngine:
        fp = fingerprint(self)
        if self._engine is None or fp != getattr(self, "_engine_fp", None):
            head = self.heads["map"]
            eng = BevDecodeEngine(timesteps=self.timesteps, time_difference=self.time_difference,
                                  noise_schedule=self.noise_schedule, diffusion=self.diffusion, bit_scale=self.bit_scale,
                                  threshold=self.threshold, feat_channels=self.feat_channels,
                                  learned_sinusoidal_dim=self.learned_sinusoidal_dim, num_layers=head.encoder.num_layers,
                                  sample_range=self.sample_range, gemm_mode=self.gemm_mode)
            eng.load_state_dict(self.state_dict())
            self._engine, self._engine_fp = eng, fp
        return self._engine

    def refresh_engine(self):
        self._engine = None

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.refresh_engine()
        return out

    def _apply(self, fn, *a, **k):
        self.refresh_engine()
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def ddim_sample(self, x, head=None, noise=None):
        """x: [fused BEV feature (b, feat_channels, h, w)] as in the reference (:269-275).  The reference loop is defined
        for b = 1; here every image gets its own ``randsteps`` samples and its own mean.  -> (b, 6, H', W')."""
        head = head if head is not None else self.heads["map"]
        feat = x[0] if isinstance(x, (list, tuple)) else x
        b, c, h, w = feat.shape
        if noise is None:
            noise = torch.randn((b, self.randsteps, 256, h, w), device=feat.device)      # :275
        gy, gx = head.grid_coords()
        return self.engine().sample(feat, noise, gy, gx)

    def ddpm_sample(self, x, head=None):
        raise NotImplementedError("the BEV reference's ddpm_sample (:303-342) cannot run as written; only ddim is built")
