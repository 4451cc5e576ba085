"""The neck in front of the decode loop: ``FPN`` + ``MultiStageMerging`` plug-ins over the C ABI.

Reference: segmentation/mmseg/models/necks/fpn.py:13-213, multi_stage_merging.py:11-52 (depth/depth/models/necks/ holds
copies).  Every DDP config chains the two (``neck=[dict(type='FPN', ...), dict(type='MultiStageMerging', ...)]``); here the
pair runs as ONE launch sequence in libddp_b200.so (``ddp_neck_forward``) when it is built from such a list
(``FusedNeck``), and each class also works on its own with the reference's forward signature.  Same constructor
arguments and state-dict keys as the reference; arguments no DDP config uses are rejected with NotImplementedError
instead of being silently approximated.  No tensor math happens in Python and there is no CPU fallback.
"""
import ctypes
from typing import Mapping, Sequence

import torch
import torch.nn as nn

from . import _lib as L
from ._stale import fingerprint
from .registry import NECKS

OUT = 256


class NeckEngine:
    """Python handle of one ``ddp_neck`` (FPN, MultiStageMerging or the fused pair) bound to one CUDA device."""

    def __init__(self, in_channels: Sequence[int], stages=L.NECK_STAGE_FPN | L.NECK_STAGE_MERGE, out_channels=OUT,
                 num_groups=32, eps=1e-5, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("ddp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if not 1 <= len(in_channels) <= 4:
            raise ValueError(f"the neck takes 1..4 levels, got {len(in_channels)}")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self.in_channels = [int(c) for c in in_channels]
        self.stages = int(stages)
        self.levels = len(in_channels)
        cfg = L.DDPNeckConfig(abi_version=L.ABI_VERSION, stages=self.stages, num_levels=self.levels,
                              in_channels=(ctypes.c_int32 * 4)(*(self.in_channels + [0] * (4 - self.levels))),
                              out_channels=out_channels, num_groups=num_groups, eps=eps)
        self._h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.ddp_neck_create(ctypes.byref(cfg), ctypes.byref(self._h))
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_neck_last_error(None).decode())
        self._plan = None
        self._ws = None

    def _check(self, rc):
        if rc != 0:
            raise L.DDPError(rc, self.lib.ddp_neck_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.ddp_neck_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_names(self):
        out, n = {}, ctypes.c_int64()
        for i in range(self.lib.ddp_neck_weight_count(self._h)):
            # two statements on purpose: `out[f(byref(n))] = n.value` reads n.value BEFORE the call (Python evaluates the
            # right-hand side first) — the round-1 bug that failed every neck / BEV hardware test
            name = self.lib.ddp_neck_weight_name(self._h, i, ctypes.byref(n)).decode()
            out[name] = n.value
        return out

    def load_state_dict(self, sd: Mapping[str, torch.Tensor]):
        """Take the neck entries of a state dict; keys may carry a module path in front
        (``neck.0.lateral_convs.0.conv.weight`` / ``lateral_convs.0.conv.weight``)."""
        by_stem = {}
        for k, v in sd.items():
            for stem in ("lateral_convs.", "fpn_convs.", "down."):
                i = k.find(stem)
                if i >= 0:
                    by_stem[k[i:]] = v
                    break
        need = self.weight_names()
        missing = [k for k in need if k not in by_stem]
        if missing:
            raise KeyError(f"state dict lacks neck weights: {missing[:4]}{'...' if len(missing) > 4 else ''}")
        for k, numel in need.items():
            t = by_stem[k].detach().to("cpu", torch.float32).contiguous()
            if t.numel() != numel:
                raise ValueError(f"{k}: expected {numel} elements, got {tuple(t.shape)}")
            self._check(self.lib.ddp_neck_set_weight(self._h, k.encode(), ctypes.c_void_p(t.data_ptr()), numel))
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_neck_commit_weights(self._h))

    def plan(self, B, sizes):
        key = (B, tuple(sizes))
        if self._plan == key:
            return
        hs = (ctypes.c_int32 * self.levels)(*[s[0] for s in sizes])
        ws = (ctypes.c_int32 * self.levels)(*[s[1] for s in sizes])
        nbytes = ctypes.c_size_t()
        self._check(self.lib.ddp_neck_plan(self._h, B, hs, ws, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value + 256:
            self._ws = torch.empty(nbytes.value + 256, dtype=torch.uint8, device=self.device)
        self._ws_bytes = nbytes.value
        self._plan = key

    def _ws_ptr(self):
        return (self._ws.data_ptr() + 255) // 256 * 256          # the library wants a 256-byte aligned workspace

    @property
    def last_launch_count(self):
        return int(self.lib.ddp_neck_last_launch_count(self._h))

    def forward(self, inputs: Sequence[torch.Tensor], want_fpn=False):
        """inputs: the pyramid, NCHW fp32 CUDA tensors.  Returns (x or None, fpn_outs or None)."""
        if len(inputs) != self.levels:
            raise AssertionError(f"expected {self.levels} inputs, got {len(inputs)}")          # fpn.py:164 / msm.py:41
        merge = bool(self.stages & L.NECK_STAGE_MERGE)
        want_fpn = want_fpn or not merge
        xs = []
        for l, x in enumerate(inputs):
            c = self.in_channels[l] if self.stages & L.NECK_STAGE_FPN else OUT
            if x.dim() != 4 or x.shape[1] != c or x.shape[0] != inputs[0].shape[0]:
                raise ValueError(f"inputs[{l}] has shape {tuple(x.shape)}, expected (B, {c}, h, w)")
            if x.device != self.device:
                raise ValueError(f"inputs[{l}] is on {x.device}, the neck engine is on {self.device}")
            xs.append(x.detach().to(torch.float32).contiguous())
        B = xs[0].shape[0]
        sizes = [tuple(x.shape[2:]) for x in xs]
        if B == 0:
            empty = [xs[0].new_empty((0, OUT) + s) for s in sizes]
            return (empty[0] if merge else None), (empty if want_fpn else None)
        self.plan(B, sizes)
        x_out = torch.empty((B, OUT) + sizes[0], dtype=torch.float32, device=self.device) if merge else None
        fpn = [torch.empty((B, OUT) + s, dtype=torch.float32, device=self.device) for s in sizes] if want_fpn else None
        vp = ctypes.c_void_p
        in_ptrs = (vp * self.levels)(*[x.data_ptr() for x in xs])
        fpn_ptrs = (vp * self.levels)(*[t.data_ptr() for t in fpn]) if fpn else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            self._check(self.lib.ddp_neck_forward(self._h, in_ptrs, vp(x_out.data_ptr()) if merge else None, fpn_ptrs,
                                                  vp(self._ws_ptr()), self._ws_bytes, vp(stream)))
        return x_out, fpn


# ---- parameter containers with the reference's state-dict keys ------------------------------------------------------
class _ConvGN(nn.Module):
    """mmcv ConvModule(conv -> GN, no conv bias, no activation): keys ``conv.weight``, ``gn.weight``, ``gn.bias``."""

    def __init__(self, cin, cout, k, num_groups):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False)
        self.gn = nn.GroupNorm(num_groups, cout)
        nn.init.xavier_uniform_(self.conv.weight)          # init_cfg=dict(type='Xavier', layer='Conv2d', distribution='uniform')


def _gn_groups(norm_cfg, who):
    if not isinstance(norm_cfg, Mapping) or norm_cfg.get("type") != "GN":
        raise NotImplementedError(f"{who}: ddp_b200 builds the neck the DDP configs use (norm_cfg=dict(type='GN', "
                                  f"num_groups=32)); got norm_cfg={norm_cfg!r}")
    return int(norm_cfg.get("num_groups", 32))


class _NeckModule(nn.Module):
    """Engine management shared by the three neck modules (rebuilt lazily after parameters change or move)."""

    _stages = 0

    def _engine_state(self):
        return self.state_dict()

    def engine(self) -> NeckEngine:
        fp = fingerprint(self)
        if self._engine is None or fp != getattr(self, "_engine_fp", None):
            eng = NeckEngine(self._engine_in_channels(), stages=self._stages, num_groups=self._groups)
            eng.load_state_dict(self._engine_state())
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

    def init_weights(self):
        pass


@NECKS.register_module()
class FPN(_NeckModule):
    """segmentation/mmseg/models/necks/fpn.py:13-213 for the argument set of the DDP configs."""

    _stages = L.NECK_STAGE_FPN

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 extra_convs_on_inputs=False, relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None,
                 norm_cfg=None, act_cfg=None, upsample_cfg=dict(mode="nearest"), init_cfg=None):
        super().__init__()
        assert isinstance(in_channels, (list, tuple))                                           # fpn.py:84
        self.in_channels = list(in_channels)
        self.out_channels = out_channels
        self.num_ins = len(in_channels)
        self.num_outs = num_outs
        unsupported = dict(start_level=start_level != 0, end_level=end_level not in (-1, self.num_ins),
                           add_extra_convs=bool(add_extra_convs), no_norm_on_lateral=no_norm_on_lateral,
                           conv_cfg=conv_cfg is not None, act_cfg=act_cfg is not None,
                           upsample_cfg=dict(upsample_cfg) != dict(mode="nearest"), num_outs=num_outs != self.num_ins,
                           out_channels=out_channels != OUT)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"FPN: ddp_b200 builds the FPN the DDP configs use; unsupported arguments: {bad}")
        self._groups = _gn_groups(norm_cfg, "FPN")
        self.lateral_convs = nn.ModuleList(_ConvGN(c, out_channels, 1, self._groups) for c in in_channels)
        self.fpn_convs = nn.ModuleList(_ConvGN(out_channels, out_channels, 3, self._groups) for _ in in_channels)
        self._engine = None

    def _engine_in_channels(self):
        return self.in_channels

    @torch.no_grad()
    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)                                             # fpn.py:164
        _, outs = self.engine().forward(list(inputs))
        return tuple(outs)


@NECKS.register_module()
class MultiStageMerging(_NeckModule):
    """segmentation/mmseg/models/necks/multi_stage_merging.py:11-52 for the argument set of the DDP configs."""

    _stages = L.NECK_STAGE_MERGE

    def __init__(self, in_channels, out_channels, kernel_size=1, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 align_corners=False, init_cfg=None):
        super().__init__()
        assert isinstance(in_channels, (list, tuple))                                           # msm.py:25
        self.in_channels = list(in_channels)
        self.out_channels = out_channels
        self.align_corners = align_corners
        unsupported = dict(kernel_size=kernel_size != 1, conv_cfg=conv_cfg is not None, act_cfg=act_cfg is not None,
                           align_corners=bool(align_corners), out_channels=out_channels != OUT,
                           in_channels=any(c != OUT for c in in_channels))
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"MultiStageMerging: ddp_b200 builds the merge the DDP configs use; unsupported arguments: {bad}")
        self._groups = _gn_groups(norm_cfg, "MultiStageMerging")
        self.down = _ConvGN(sum(in_channels), out_channels, 1, self._groups)
        self._engine = None

    def _engine_in_channels(self):
        return self.in_channels

    @torch.no_grad()
    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)                                             # msm.py:41
        x, _ = self.engine().forward(list(inputs))
        return [x]


class FusedNeck(nn.Sequential, _NeckModule):
    """``neck=[FPN, MultiStageMerging]`` as ONE library call.  Children keep the indices 0 / 1, so the state-dict keys
    are the reference's (``neck.0.lateral_convs...``, ``neck.1.down...``); ``self[0](...)`` / ``self[1](...)`` still
    run the single stages."""

    _stages = L.NECK_STAGE_FPN | L.NECK_STAGE_MERGE

    def __init__(self, fpn: FPN, msm: MultiStageMerging):
        nn.Sequential.__init__(self, fpn, msm)
        if len(msm.in_channels) != fpn.num_outs or fpn._groups != msm._groups:
            raise ValueError("FusedNeck: MultiStageMerging does not match the FPN in front of it")
        self._groups = fpn._groups
        self._engine = None

    def _engine_in_channels(self):
        return self[0].in_channels

    @torch.no_grad()
    def forward(self, inputs):
        assert len(inputs) == len(self[0].in_channels)
        x, _ = self.engine().forward(list(inputs))
        return [x]


def fuse_neck(modules):
    """[FPN, MultiStageMerging] -> FusedNeck; anything else stays a plain nn.Sequential."""
    if len(modules) == 2 and isinstance(modules[0], FPN) and isinstance(modules[1], MultiStageMerging):
        return FusedNeck(modules[0], modules[1])
    return nn.Sequential(*modules)
