"""Minimal stand-in for the mmcv Registry the reference plugs into (mmcv is not a dependency here).

The reference registers its classes with ``@SEGMENTORS.register_module()`` / ``@HEADS.register_module()``
(segmentation/mmseg/models/builder.py:8-15: SEGMENTORS is HEADS is MODELS) and builds them from config
dicts keyed by ``type``.  ``register_into_mmseg()`` puts the same classes into a real mmseg/depth registry
(``force=True``) when those packages are importable, which is how the drop-in is used inside the reference.
"""
import copy
import warnings


class Registry:
    def __init__(self, name):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            key = name or cls.__name__
            if key in self._modules and not force:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._modules[key] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self._modules.get(key)

    def __contains__(self, key):
        return key in self._modules

    def build(self, cfg, **default_args):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"cfg must be a dict with a `type` key, got {cfg!r}")
        args = copy.deepcopy(dict(cfg))
        typ = args.pop("type")
        cls = typ if isinstance(typ, type) else self.get(typ)
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry")
        for k, v in default_args.items():
            args.setdefault(k, v)
        try:
            return cls(**args)
        except Exception as e:               # same message shape as mmcv.utils.build_from_cfg
            raise type(e)(f"{cls.__name__}: {e}")


MODELS = Registry("models")
BACKBONES = NECKS = HEADS = LOSSES = SEGMENTORS = MODELS       # one registry, as in mmseg/models/builder.py
DEPTHER = Registry("depther")                                  # depth/depth/models/builder.py keeps a separate one


def build_backbone(cfg):
    return MODELS.build(cfg)


def build_neck(cfg):
    return MODELS.build(cfg)


def build_head(cfg):
    return MODELS.build(cfg)


def build_segmentor(cfg, train_cfg=None, test_cfg=None):
    """segmentation/mmseg/models/builder.py:38-49."""
    if train_cfg is not None or test_cfg is not None:
        warnings.warn("train_cfg and test_cfg is deprecated, please specify them in model", UserWarning)
    assert cfg.get("train_cfg") is None or train_cfg is None, "train_cfg specified in both outer field and model field"
    assert cfg.get("test_cfg") is None or test_cfg is None, "test_cfg specified in both outer field and model field"
    return SEGMENTORS.build(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


def build_depther(cfg, train_cfg=None, test_cfg=None):
    return DEPTHER.build(cfg, train_cfg=train_cfg, test_cfg=test_cfg)


def host_framework_builder(typ):
    """A class named `typ` from the host framework's own registries (real mmseg / depth toolbox), or None.
    Used for the parts of a config that are outside ddp_b200's scope (backbones, stock necks) when the plug-ins run
    inside the reference: the reference's registries build them, exactly as they would without ddp_b200."""
    for mod in ("mmseg.models.builder", "depth.models.builder"):
        try:
            m = __import__(mod, fromlist=["BACKBONES", "NECKS"])
        except ImportError:
            continue
        for reg in (getattr(m, "BACKBONES", None), getattr(m, "NECKS", None)):
            if reg is not None and reg.get(typ) is not None:
                return reg
    return None


def register_into_mmseg(necks=False):
    """Register the B200 classes into the real mmseg / depth / mmdet3d registries (force=True).  Returns what was done.

    The DDP segmentors build ``neck=[FPN, MultiStageMerging]`` into the fused CUDA neck by themselves; ``necks=True``
    additionally REPLACES the stock ``FPN`` / ``MultiStageMerging`` entries of the host registries (only do that in a
    process that runs nothing but DDP configs: the replacements refuse arguments those configs do not use)."""
    done = []
    from .models import ddp as seg_mod, depth_ddp as depth_mod, deformable_head_with_time as head_mod
    from . import neck as neck_mod, bev as bev_mod
    try:
        from mmseg.models.builder import SEGMENTORS as S, HEADS as H, NECKS as N
        S.register_module(name="DDP", force=True, module=seg_mod.DDP)
        S.register_module(name="SelfAlignedDDP", force=True, module=seg_mod.SelfAlignedDDP)
        H.register_module(name="DeformableHeadWithTime", force=True, module=head_mod.DeformableHeadWithTime)
        if necks:
            N.register_module(name="FPN", force=True, module=neck_mod.FPN)
            N.register_module(name="MultiStageMerging", force=True, module=neck_mod.MultiStageMerging)
        done.append("mmseg")
    except ImportError:
        pass
    try:
        from depth.models.builder import DEPTHER as D, HEADS as DH, NECKS as DN
        D.register_module(name="DDP", force=True, module=depth_mod.DDP)
        DH.register_module(name="DeformableHeadWithTime", force=True, module=head_mod.DepthDeformableHeadWithTime)
        if necks:
            DN.register_module(name="FPN", force=True, module=neck_mod.FPN)
            DN.register_module(name="MultiStageMerging", force=True, module=neck_mod.MultiStageMerging)
        done.append("depth")
    except ImportError:
        pass
    try:        # bev/mmdet3d/models/builder.py: FUSIONMODELS / HEADS
        from mmdet3d.models.builder import FUSIONMODELS as F3, HEADS as H3
        F3.register_module(name="DDP", force=True, module=bev_mod.BevDDP)
        H3.register_module(name="DeformableHeadWithTime", force=True, module=bev_mod.BevDeformableHeadWithTime)
        done.append("mmdet3d")
    except ImportError:
        pass
    return done
