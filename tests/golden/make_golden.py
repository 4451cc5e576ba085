"""Generate golden fixtures by running the UNMODIFIED reference here.

    python tests/golden/make_golden.py seg
    python tests/golden/make_golden.py depth
    python tests/golden/make_golden.py neck
    python tests/golden/make_golden.py bev

Runs only in the build container (needs /root/reference).  It builds the
reference's own classes from the reference's own config files
(``segmentation/configs/...``, ``depth/configs/...``) through the import shim
in ``refshim.py``, loads the seeded weights of ``oracle.ddp_oracle.make_weights``
into them by state-dict key, calls ``DDP.ddim_sample`` / ``DDP.sample`` on
seeded inputs and stores outputs (+ per-step head outputs) as ``.npz``.

The fixtures hold seeds, shapes, checksums of the inputs and the reference's
outputs; ``tests/test_oracle_golden.py`` regenerates inputs from the seeds and
compares the oracle with the stored outputs.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import refshim  # noqa: E402
from oracle import ddp_oracle as O  # noqa: E402
from oracle import neck_oracle as NO  # noqa: E402
from oracle import bev_oracle as BO  # noqa: E402

SEG_CASES = [
    # name, config file, class override, num_classes, T, R, accumulation, h, w, wseed, xseed
    dict(name="seg_city_T3", cfg="cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py",
         T=None, R=1, h=12, w=20, wseed=11, xseed=101),
    dict(name="seg_ade_T3_acc", cfg="ade/ddp_swin_t_2x8_512x512_160k_ade20k.py",
         T=None, R=1, h=16, w=16, wseed=12, xseed=102),
    dict(name="seg_aligned_T10_R2", cfg="cityscapes/ddp_convnext_t_4x4_512x1024_5k_cityscapes_aligned.py",
         T=None, R=2, h=8, w=16, wseed=13, xseed=103),
    dict(name="seg_plumbing_T1_64x64", cfg="cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py",
         T=1, R=1, h=64, w=64, wseed=14, xseed=104),
    dict(name="seg_city_T3_R3_ragged", cfg="cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py",
         T=None, R=3, h=7, w=13, wseed=15, xseed=105),
    # the stochastic sampler (ddp.py:248-290): diffusion='ddpm' is a constructor argument of the reference class
    dict(name="seg_ddpm_T4_R2", cfg="cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py",
         T=4, R=2, h=10, w=12, wseed=16, xseed=106, diffusion="ddpm"),
]
DEPTH_CASES = [
    dict(name="depth_nyu_T3", cfg="ddp_nyu/ddp_swint_1k_w7_nyu_bs2x8_scale01.py",
         T=None, R=1, h=12, w=16, wseed=21, xseed=201),
    dict(name="depth_nyu_T20_R2", cfg="ddp_nyu/ddp_swint_1k_w7_nyu_bs2x8_scale01.py",
         T=20, R=2, h=9, w=11, wseed=22, xseed=202),
]

# the neck in front of the loop: FPN + MultiStageMerging built from the reference's config files
NECK_CASES = [
    dict(name="neck_swin_t_city", tree="segmentation", cfg="cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py",
         B=1, h=12, w=20, wseed=31, xseed=301),
    dict(name="neck_swin_l_ade", tree="segmentation", cfg="ade/ddp_swin_l_2x8_512x512_160k_ade20k.py",
         B=2, h=8, w=8, wseed=32, xseed=302),
    # odd pyramid sizes (15x20 -> 8x10 -> 4x5 -> 2x3): nearest / bilinear ratios that are not exactly 2
    dict(name="neck_convnext_t_odd", tree="segmentation",
         cfg="cityscapes/ddp_convnext_t_4x4_512x1024_5k_cityscapes_aligned.py", B=1, h=15, w=20, wseed=33, xseed=303),
]


def checksum(t):
    t = t.double()
    return np.array([float(t.sum()), float(t.abs().sum())])


def load_weights(model, W):
    sd = model.state_dict()
    hot = [k for k in sd if not k.startswith(("backbone.", "neck.", "auxiliary_head."))]
    missing = [k for k in hot if k not in W]
    extra = [k for k in W if k not in sd]
    assert not missing and not extra, (missing, extra)
    for k in hot:
        assert tuple(sd[k].shape) == tuple(W[k].shape), (k, sd[k].shape, W[k].shape)
    model.load_state_dict(W, strict=False)


def record_head(model, store):
    orig = model._decode_head_forward_test

    def wrapped(x, t, img_metas):
        out = orig(x, t, img_metas)
        store.append(out.detach().clone())
        return out
    model._decode_head_forward_test = wrapped


def run_seg():
    refshim.install("segmentation")
    # the aligned configs do custom_imports='mmcls.models' (absent here; only provides the backbone)
    for n in ("mmcls", "mmcls.models"):
        sys.modules[n] = types.ModuleType(n)
    import mmcv  # noqa: F401
    from mmcv import Config
    # mmcls.ConvNeXt (aligned configs) is not importable here; the backbone never runs in ddim_sample.
    from mmseg.models import build_segmentor
    from mmseg.models.builder import BACKBONES

    @BACKBONES.register_module(name="NoBackbone")
    class _NoBackbone(torch.nn.Module):
        def __init__(self, **kw):
            super().__init__()

        def init_weights(self):
            pass

    for case in SEG_CASES:
        cfg = Config.fromfile(f"{refshim.REF}/segmentation/configs/{case['cfg']}")
        m = cfg.model
        m.pretrained = None
        if m.backbone.type.startswith("mmcls."):
            m.backbone = dict(type="NoBackbone")       # scoped mmcls type is unresolvable here
        if "init_cfg" in m.backbone:
            m.backbone.init_cfg = None
        m.auxiliary_head.norm_cfg = dict(type="BN")   # SyncBN -> BN, as tools/test.py:278 does
        m.decode_head.norm_cfg = dict(type="BN")
        if case["T"] is not None:
            m.timesteps = case["T"]
        m.randsteps = case["R"]
        diffusion = case.get("diffusion", "ddim")
        m.diffusion = diffusion
        model = build_segmentor(m)
        model.eval()
        ocfg = O.OracleConfig(task="seg", num_classes=m.decode_head.num_classes, timesteps=m.timesteps,
                              randsteps=case["R"], bit_scale=m.bit_scale,
                              accumulation=bool(m.get("accumulation", False)), diffusion=diffusion)
        W = O.make_weights(ocfg, seed=case["wseed"])
        load_weights(model, W)
        h, w, R = case["h"], case["w"], case["R"]
        x = torch.randn(1, 256, h, w, generator=torch.Generator().manual_seed(case["xseed"]))
        nseed = case["xseed"] + 1000
        torch.manual_seed(nseed)
        noise = torch.randn((R, 256, h, w))            # what ddp.py:220 will draw
        steps = []
        record_head(model, steps)
        torch.manual_seed(nseed)
        out = model.ddim_sample(x, None) if diffusion == "ddim" else model.ddpm_sample(x, None)
        np.savez_compressed(
            os.path.join(HERE, case["name"] + ".npz"),
            task="seg", model_type=m.type, diffusion=diffusion, config=case["cfg"], num_classes=ocfg.num_classes,
            timesteps=ocfg.timesteps, randsteps=R, bit_scale=ocfg.bit_scale,
            accumulation=ocfg.accumulation, h=h, w=w, wseed=case["wseed"], xseed=case["xseed"],
            nseed=nseed, x_checksum=checksum(x), noise_checksum=checksum(noise),
            w_checksum=checksum(torch.cat([v.flatten() for _, v in sorted(W.items())])),
            out=out.numpy(), step_logits=torch.stack(steps).numpy().astype(np.float32),
            step_argmax=torch.stack(steps).argmax(2).numpy().astype(np.int16))
        print(case["name"], m.type, tuple(out.shape), float(out.abs().mean()))


def run_depth():
    refshim.install("depth")
    for n in ("timm", "timm.models", "timm.models.layers", "matplotlib", "matplotlib.pyplot",
              "matplotlib.cm", "matplotlib.colors"):
        class _Any(types.ModuleType):
            def __getattr__(self, k):
                if k.startswith("__"):
                    raise AttributeError(k)
                return _Any(self.__name__ + "." + k)

            def __call__(self, *a, **k):
                return None
        sys.modules[n] = _Any(n)
    import mmcv  # noqa: F401
    # depth/depth/models/depther/__init__.py imports a module that is not in the tree
    # (regulardepth); load the two files we need without running that __init__.
    dp = types.ModuleType("depth.models.depther")
    dp.__path__ = [f"{refshim.REF}/depth/depth/models/depther"]
    sys.modules["depth.models.depther"] = dp
    from depth.models import build_depther
    import depth.models.depther.ddp  # noqa: F401
    # The depth tree has no time-aware BaseTransformerLayer (it relies on a
    # patched site-packages mmcv); the segmentation fork is the in-tree statement of that layer.
    import importlib.util
    from mmcv.utils.registry import Registry
    orig = Registry._register_module

    def forced(self, module_class, module_name=None, force=False):
        return orig(self, module_class, module_name, True)
    Registry._register_module = forced
    spec = importlib.util.spec_from_file_location(
        "seg_transformer_fork", f"{refshim.REF}/segmentation/mmseg/models/utils/transformer.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    Registry._register_module = orig
    from mmcv import Config

    for case in DEPTH_CASES:
        cfg = Config.fromfile(f"{refshim.REF}/depth/configs/{case['cfg']}")
        m = cfg.model
        m.backbone.init_cfg = None
        if case["T"] is not None:
            m.timesteps = case["T"]
        m.randsteps = case["R"]
        model = build_depther(m)
        model.eval()
        ocfg = O.OracleConfig(task="depth", timesteps=m.timesteps, randsteps=case["R"],
                              bit_scale=m.bit_scale, min_depth=m.min_depth, max_depth=m.max_depth)
        W = O.make_weights(ocfg, seed=case["wseed"])
        load_weights(model, W)
        h, w, R = case["h"], case["w"], case["R"]
        x = torch.randn(1, 256, h, w, generator=torch.Generator().manual_seed(case["xseed"]))
        nseed = case["xseed"] + 1000
        torch.manual_seed(nseed)
        noise = torch.randn((R, 1, h, w))
        steps = []
        record_head(model, steps)
        torch.manual_seed(nseed)
        out = model.sample(x, None)
        out = torch.clamp(out, min=model.decode_head.min_depth, max=model.decode_head.max_depth)
        np.savez_compressed(
            os.path.join(HERE, case["name"] + ".npz"),
            task="depth", model_type=m.type, config=case["cfg"], timesteps=ocfg.timesteps, randsteps=R,
            bit_scale=ocfg.bit_scale, min_depth=ocfg.min_depth, max_depth=ocfg.max_depth,
            h=h, w=w, wseed=case["wseed"], xseed=case["xseed"], nseed=nseed,
            x_checksum=checksum(x), noise_checksum=checksum(noise),
            w_checksum=checksum(torch.cat([v.flatten() for _, v in sorted(W.items())])),
            out=out.numpy(), step_depth=torch.stack(steps).numpy().astype(np.float32))
        print(case["name"], tuple(out.shape), float(out.mean()), float(out.min()), float(out.max()))


def run_neck():
    refshim.install("segmentation")
    for n in ("mmcls", "mmcls.models"):
        sys.modules[n] = types.ModuleType(n)
    import mmcv  # noqa: F401
    from mmcv import Config
    import mmseg.models  # noqa: F401  (registers FPN / MultiStageMerging)
    from mmseg.models import builder

    for case in NECK_CASES:
        cfg = Config.fromfile(f"{refshim.REF}/{case['tree']}/configs/{case['cfg']}")
        ncfg = cfg.model.neck
        neck = builder.build_neck(ncfg)                      # Sequential(FPN, MultiStageMerging), neck.0.* / neck.1.*
        neck.eval()
        in_channels = list(ncfg[0].in_channels)
        W = NO.make_weights(in_channels, seed=case["wseed"])
        sd = neck.state_dict()
        assert sorted("neck." + k for k in sd) == sorted(W), (sorted(sd), sorted(W))
        for k in sd:
            assert tuple(sd[k].shape) == tuple(W["neck." + k].shape), k
        neck.load_state_dict({k[len("neck."):]: v for k, v in W.items()})
        xs = NO.make_inputs(in_channels, case["B"], case["h"], case["w"], seed=case["xseed"])
        with torch.no_grad():
            fpn_outs = neck[0](tuple(xs))
            out = neck[1](fpn_outs)
            assert isinstance(out, list) and len(out) == 1
            whole = neck(tuple(xs))[0]
        assert torch.equal(whole, out[0])
        np.savez_compressed(
            os.path.join(HERE, case["name"] + ".npz"),
            task="neck", config=case["cfg"], in_channels=np.array(in_channels), B=case["B"], h=case["h"], w=case["w"],
            wseed=case["wseed"], xseed=case["xseed"],
            x_checksum=checksum(torch.cat([x.flatten() for x in xs])),
            w_checksum=checksum(torch.cat([v.flatten() for _, v in sorted(W.items())])),
            out=out[0].numpy(), **{f"fpn{i}": o.numpy() for i, o in enumerate(fpn_outs)})
        print(case["name"], in_channels, tuple(out[0].shape), float(out[0].abs().mean()))


# BEV map segmentation.  Scopes chosen so that the state grid is small: same structure as the shipped
# grid_transform (input step 0.8, output step 0.5, output range inside the input range), fewer cells.
BEV_CASES = [
    dict(name="bev_camera_T3_R5", cfg="nuscenes/seg/ddp-camera-bev256d2-lss-scale001-d5-lr5e-5.yaml", T=None, R=None,
         input_scope=((-6.4, 6.4, 0.8), (-8.0, 8.0, 0.8)), output_scope=((-6.0, 6.0, 0.5), (-7.5, 7.5, 0.5)),
         wseed=41, xseed=401),
    dict(name="bev_fusion_T2_R2", cfg="nuscenes/seg/ddp-fusion-bev256d2-lss-scale001-d5-lr5e-5.yaml", T=2, R=2,
         input_scope=((-4.8, 4.8, 0.8), (-4.0, 4.0, 0.8)), output_scope=((-4.5, 4.5, 0.5), (-3.5, 3.5, 0.5)),
         wseed=42, xseed=402),
]


def run_bev():
    """bev/mmdet3d is not importable as a package here (its __init__ pulls compiled ops); load the two hot-path files
    by path with the package context stubbed, exactly as written."""
    import importlib.util
    import yaml
    refshim.install("segmentation")            # mmcv alias + the time-aware transformer layer fork
    import mmcv  # noqa: F401
    import mmseg.models  # noqa: F401  registers BaseTransformerLayer(time), DetrTransformerEncoder, SinePositionalEncoding
    from mmcv.utils import Registry
    bev = f"{refshim.REF}/bev"

    def pkg(name, path=None):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
        return m
    pkg("mmdet3d", f"{bev}/mmdet3d")
    models = pkg("mmdet3d.models", f"{bev}/mmdet3d/models")
    builder = types.ModuleType("mmdet3d.models.builder")
    for n in ("build_backbone", "build_fuser", "build_head", "build_neck", "build_vtransform"):
        setattr(builder, n, lambda *a, **k: None)
    builder.HEADS = Registry("bev_heads")
    builder.FUSIONMODELS = Registry("bev_fusion_models")
    sys.modules["mmdet3d.models.builder"] = builder
    models.FUSIONMODELS = builder.FUSIONMODELS
    ops = pkg("mmdet3d.ops", f"{bev}/mmdet3d/ops")
    ops.Voxelization = ops.DynamicScatter = object
    spec = importlib.util.spec_from_file_location("mmdet3d.ops.norm", f"{bev}/mmdet3d/ops/norm.py")
    norm = importlib.util.module_from_spec(spec)
    sys.modules["mmdet3d.ops.norm"] = norm
    spec.loader.exec_module(norm)               # the reference's own `resize`
    pkg("mmdet3d.models.fusion_models", f"{bev}/mmdet3d/models/fusion_models")
    bf = types.ModuleType("mmdet3d.models.fusion_models.bevfusion")

    class BEVFusion(torch.nn.Module):           # the encoder side of the model never runs in ddim_sample
        def __init__(self, **kw):
            super().__init__()
    bf.BEVFusion = BEVFusion
    sys.modules["mmdet3d.models.fusion_models.bevfusion"] = bf

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    ddp_mod = load("mmdet3d.models.fusion_models.ddp", f"{bev}/mmdet3d/models/fusion_models/ddp.py")
    head_mod = load("mmdet3d.models.heads.segm.deformable_head_with_time",
                    f"{bev}/mmdet3d/models/heads/segm/deformable_head_with_time.py")
    base = yaml.safe_load(open(f"{bev}/configs/nuscenes/default.yaml"))
    for case in BEV_CASES:
        y = yaml.safe_load(open(f"{bev}/configs/{case['cfg']}"))["model"]
        from mmcv.utils import ConfigDict
        hc = ConfigDict(y["heads"]["map"])                   # attribute access, as mmcv's Config gives the head
        assert hc.pop("type") == "DeformableHeadWithTime"
        T = case["T"] or y["timesteps"]
        R = case["R"] or y["randsteps"]
        head = head_mod.DeformableHeadWithTime(
            classes=base["map_classes"], loss="focal",
            grid_transform=dict(input_scope=[list(s) for s in case["input_scope"]],
                                output_scope=[list(s) for s in case["output_scope"]]), **hc)
        model = ddp_mod.DDP(bit_scale=y["bit_scale"], timesteps=T, randsteps=R, time_difference=y["time_difference"],
                            learned_sinusoidal_dim=y["learned_sinusoidal_dim"], sample_range=tuple(y["sample_range"]),
                            noise_schedule=y["noise_schedule"], diffusion=y["diffusion"],
                            **({"feat_channels": y["feat_channels"]} if "feat_channels" in y else {}))
        feat = model.transform.conv.in_channels - 256
        model.heads = torch.nn.ModuleDict({"map": head})
        model.eval()
        bcfg = BO.BevConfig(timesteps=T, randsteps=R, time_difference=y["time_difference"], bit_scale=y["bit_scale"],
                            num_layers=hc["encoder"]["num_layers"], feat_channels=feat,
                            input_scope=case["input_scope"], output_scope=case["output_scope"])
        W = BO.make_weights(bcfg, seed=case["wseed"])
        sd = model.state_dict()
        assert sorted(sd) == sorted(W), (sorted(set(sd) ^ set(W)))
        for k in sd:
            assert tuple(sd[k].shape) == tuple(W[k].shape), k
        model.load_state_dict(W)
        h = len(torch.arange(*case["input_scope"][0]))
        w = len(torch.arange(*case["input_scope"][1]))
        x = torch.randn(1, feat, h, w, generator=torch.Generator().manual_seed(case["xseed"]))
        nseed = case["xseed"] + 1000
        torch.manual_seed(nseed)
        noise = torch.randn((R, 256, h, w))              # what fusion_models/ddp.py:275 will draw
        steps = []
        orig = head.forward

        def wrapped(inputs, times, target=None, _orig=orig, _steps=steps):
            out = _orig(inputs, times, target)
            _steps.append(out.detach().clone())
            return out
        head.forward = wrapped
        torch.manual_seed(nseed)
        out = model.ddim_sample([x], head)
        np.savez_compressed(
            os.path.join(HERE, case["name"] + ".npz"),
            task="bev", config=case["cfg"], timesteps=T, randsteps=R, bit_scale=bcfg.bit_scale,
            num_layers=bcfg.num_layers, feat_channels=feat, input_scope=np.array(case["input_scope"]), output_scope=np.array(case["output_scope"]),
            h=h, w=w, wseed=case["wseed"], xseed=case["xseed"], nseed=nseed, x_checksum=checksum(x),
            noise_checksum=checksum(noise), w_checksum=checksum(torch.cat([v.flatten() for _, v in sorted(W.items())])),
            out=out.numpy(), step_prob=torch.stack(steps).numpy().astype(np.float32))
        print(case["name"], (h, w), tuple(out.shape), float(out.mean()), float((out > 0.5).float().mean()))


if __name__ == "__main__":
    {"seg": run_seg, "depth": run_depth, "neck": run_neck, "bev": run_bev}[sys.argv[1]]()
