"""Host-side logic of the drop-in boundary (no GPU): config loader, registry, plug-in classes' constructor
surface and state-dict keys, sharding + gather over a world_size-2 gloo group."""
import glob
import os
import warnings

import pytest
import torch

from ddp_b200.config import Config
from ddp_b200.registry import MODELS, build_segmentor, build_depther
import ddp_b200.models as M
from ddp_b200 import dist as D
from ddp_b200.neck import FPN, FusedNeck, MultiStageMerging
from oracle import ddp_oracle as O
from oracle import neck_oracle as NO

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


@MODELS.register_module(name="ToyBackbone", force=True)
class ToyBackbone(torch.nn.Module):
    """Produces a 1/stride-resolution 256-channel map (stands in for backbone + FPN + MultiStageMerging)."""

    def __init__(self, channels=256, stride=4):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, channels, stride, stride=stride)

    def forward(self, img):
        return [self.conv(img)]


def test_config_loader_base_and_delete():
    cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
    assert cfg.model.type == "DDP" and cfg.model.decode_head.encoder.num_layers == 6
    assert cfg.dist_params.backend == "nccl"                      # inherited from _base_
    assert cfg.optimizer == dict(type="AdamW", lr=0.00006, betas=(0.9, 0.999), weight_decay=0.01)   # _delete_
    assert cfg.log_config.hooks[0].type == "TextLoggerHook"


def test_build_from_config_and_state_dict_keys_match_reference():
    cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = build_segmentor(cfg.model)
    assert isinstance(model, M.DDP) and model.num_classes == 19 and model.align_corners is False
    assert model.decode_head.in_channels[0] == 256 and model.decode_head.out_channels == 19
    want = O.make_weights(O.OracleConfig(task="seg", num_classes=19), seed=0)      # the reference's own keys/shapes
    got = model._hot_state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k
    model.load_state_dict(want, strict=False)
    assert torch.equal(model.state_dict()["decode_head.encoder.layers.3.ffns.0.layers.1.weight"],
                       want["decode_head.encoder.layers.3.ffns.0.layers.1.weight"])


def test_depth_plugin_keys():
    dh = dict(type="DeformableHeadWithTime", in_channels=[256], channels=256, in_index=[0], dropout_ratio=0.,
              min_depth=1e-3, max_depth=10, num_feature_levels=1,
              encoder=dict(type="DetrTransformerEncoder", num_layers=6, transformerlayers=dict(
                  type="BaseTransformerLayer", use_time_mlp=True,
                  attn_cfgs=dict(type="MultiScaleDeformableAttention", embed_dims=256, num_levels=1, num_heads=8, dropout=0.),
                  ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024, ffn_drop=0., act_cfg=dict(type="GELU")),
                  operation_order=("self_attn", "norm", "ffn", "norm"))),
              positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5))
    model = build_depther(dict(type="DDP", bit_scale=0.1, timesteps=3, min_depth=1e-3, max_depth=10,
                               backbone=dict(type="ToyBackbone"), decode_head=dh))
    want = O.make_weights(O.OracleConfig(task="depth"), seed=0)
    got = model._hot_state_dict()
    assert set(got) == set(want)
    for k in want:
        assert tuple(got[k].shape) == tuple(want[k].shape), k


def _depth_model(test_cfg=None):
    dh = dict(type="DeformableHeadWithTime", in_channels=[256], channels=256, in_index=[0], dropout_ratio=0.,
              min_depth=1e-3, max_depth=10, num_feature_levels=1,
              encoder=dict(type="DetrTransformerEncoder", num_layers=6, transformerlayers=dict(
                  type="BaseTransformerLayer", use_time_mlp=True,
                  attn_cfgs=dict(type="MultiScaleDeformableAttention", embed_dims=256, num_levels=1, num_heads=8, dropout=0.),
                  ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024, ffn_drop=0., act_cfg=dict(type="GELU")),
                  operation_order=("self_attn", "norm", "ffn", "norm"))),
              positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5))
    return build_depther(dict(type="DDP", bit_scale=0.1, timesteps=3, min_depth=1e-3, max_depth=10,
                              backbone=dict(type="ToyBackbone"), decode_head=dh, test_cfg=test_cfg))


def test_depth_aug_test_is_the_mean_of_unflipped_views(monkeypatch):
    """depth/.../depther/encoder_decoder.py:163-229 with the shipped NYU test pipeline (MultiScaleFlipAug, flip=True: two
    views per image).  The loop itself needs the GPU; here encode_decode is replaced by a batch-independent function of
    the pixels, and aug_test (which stacks equal-size views into ONE call) must equal the reference's view-by-view
    recipe: un-flip every view's prediction, sum in view order, divide."""
    model = _depth_model(test_cfg=dict(mode="whole"))
    calls = []

    def fake_encode_decode(img, img_metas, rescale=False):
        calls.append(tuple(img.shape))
        ramp = torch.arange(img.shape[3], dtype=torch.float32).view(1, 1, 1, -1) * 0.01
        return img.mean(1, keepdim=True) * 3.0 + ramp + (0.5 if rescale else 0.0)

    monkeypatch.setattr(model, "encode_decode", fake_encode_decode)
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 12, 20, generator=g)
    meta = dict(ori_shape=(12, 20, 3), img_shape=(12, 20, 3), flip=False)
    views = [img, img.flip(3), img.flip(2), torch.randn(2, 3, 6, 10, generator=g)]
    metas = [[meta] * 2, [dict(meta, flip=True, flip_direction="horizontal")] * 2,
             [dict(meta, flip=True, flip_direction="vertical")] * 2, [dict(meta, img_shape=(6, 10, 3))] * 2]

    def reference_recipe(vs, ms):
        total = None
        for v, m in zip(vs, ms):
            p = fake_encode_decode(v, m, True)
            if m[0]["flip"]:
                p = p.flip(dims=(3,)) if m[0]["flip_direction"] == "horizontal" else p.flip(dims=(2,))
            total = p if total is None else total + p
        return total / len(vs)

    want = reference_recipe(views[:3], metas[:3])
    calls.clear()
    got = model.forward_test(views[:3], metas[:3])
    assert calls == [(6, 3, 12, 20)]                                  # three equal-size views: ONE call
    assert len(got) == 2 and all(torch.equal(torch.from_numpy(a), w) for a, w in zip(got, want))
    # the same through simple_test for one flipped view
    one = model.forward_test([views[1]], [metas[1]])
    assert torch.equal(torch.from_numpy(one[0]), fake_encode_decode(views[1], None, True).flip(3)[0])
    # batching can be switched off, and views of different sizes are never stacked
    model.test_cfg = dict(mode="whole", batch_views=False)
    calls.clear()
    got2 = model.aug_test(views[:3], metas[:3])
    assert calls == [(2, 3, 12, 20)] * 3 and all((a == b).all() for a, b in zip(got, got2))
    model.test_cfg = dict(mode="whole")
    calls.clear()
    with pytest.raises(RuntimeError):                                  # the reference's `+=` of a smaller map fails the same way
        model.aug_test(views[2:], metas[2:])
    assert calls == [(2, 3, 12, 20), (2, 3, 6, 10)]
    model.test_cfg = dict(mode="slide")
    with pytest.raises(NotImplementedError):                           # encoder_decoder.py:182-183
        model.simple_test(img, [meta] * 2)
    assert torch.equal(model.forward_dummy(img), fake_encode_decode(img, None))      # encoder_decoder.py:122-126


def test_slide_inference_equals_the_reference_window_loop(monkeypatch):
    """encoder_decoder.py:181-227 (EncoderDecoder.slide_inference), windows stacked `window_batch` at a time."""
    import torch.nn.functional as F
    cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = build_segmentor(cfg.model)
    C = model.out_channels
    wgt = torch.randn(C, 3, 1, 1, generator=torch.Generator().manual_seed(5))
    calls = []

    def fake_encode_decode(img, img_metas):
        calls.append(tuple(img.shape))
        yy = torch.arange(img.shape[2], dtype=torch.float32).view(1, 1, -1, 1) * 0.1       # depends on the position IN the window
        return F.conv2d(img, wgt) + yy

    monkeypatch.setattr(model, "encode_decode", fake_encode_decode)

    def reference_loop(img, h_crop, w_crop, h_stride, w_stride):
        b, _, h_img, w_img = img.shape
        h_grids = max(h_img - h_crop + h_stride - 1, 0) // h_stride + 1
        w_grids = max(w_img - w_crop + w_stride - 1, 0) // w_stride + 1
        preds, count = img.new_zeros((b, C, h_img, w_img)), img.new_zeros((b, 1, h_img, w_img))
        for hi in range(h_grids):
            for wi in range(w_grids):
                y1, x1 = hi * h_stride, wi * w_stride
                y2, x2 = min(y1 + h_crop, h_img), min(x1 + w_crop, w_img)
                y1, x1 = max(y2 - h_crop, 0), max(x2 - w_crop, 0)
                lg = fake_encode_decode(img[:, :, y1:y2, x1:x2], None)
                preds += F.pad(lg, (x1, w_img - x2, y1, h_img - y2))
                count[:, :, y1:y2, x1:x2] += 1
        return preds / count

    g = torch.Generator().manual_seed(9)
    for (h_img, w_img), crop, stride, wb in [((20, 28), (8, 12), (5, 7), 8), ((20, 28), (8, 12), (5, 7), 3),
                                             ((16, 16), (32, 32), (8, 8), 8), ((17, 23), 8, 6, 1)]:
        img = torch.randn(2, 3, h_img, w_img, generator=g)
        model.test_cfg = dict(mode="slide", crop_size=crop, stride=stride, window_batch=wb)
        cr, st = (crop, crop) if isinstance(crop, int) else crop, (stride, stride) if isinstance(stride, int) else stride
        want = reference_loop(img, cr[0], cr[1], st[0], st[1])
        n_windows = len(calls)
        calls.clear()
        meta = [dict(ori_shape=(h_img, w_img, 3), img_shape=(h_img, w_img, 3), flip=False)] * 2
        got = model.slide_inference(img, meta, rescale=False)
        assert torch.equal(got, want), (h_img, w_img, crop, stride)
        assert len(calls) == -(-n_windows // wb) and sum(c[0] for c in calls) == 2 * n_windows
        # through `inference`: softmax of the (rescaled) average
        out = model.inference(img, meta, rescale=True)
        assert torch.allclose(out, want.softmax(1), atol=1e-6)
        calls.clear()


def test_sampling_timesteps_and_small_surface_helpers():
    """ddp.py:198-213, depther/ddp.py:207-218, base.py:21-34: helpers the reference's loop exposes on the model."""
    cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        seg = build_segmentor(dict(cfg.model, timesteps=4, time_difference=1, sample_range=(0.1, 0.999)))
    pairs = seg._get_sampling_timesteps(3, device="cpu")
    assert len(pairs) == 4 and all(tuple(p.shape) == (2, 3) and p.dtype == torch.float32 for p in pairs)
    for step, p in enumerate(pairs):
        t_now = 1 - (step / 4) * (1 - 0.1)
        t_next = max(1 - (step + 1 + 1) / 4 * (1 - 0.1), 0.1)
        assert torch.equal(p, torch.tensor([[t_now] * 3, [t_next] * 3]))
    assert float(pairs[-1][1, 0]) == pytest.approx(0.1)                   # clamped at sample_range[0]
    t = torch.arange(3.0)
    assert seg.right_pad_dims_to(torch.zeros(3, 4, 5, 6), t).shape == (3, 1, 1, 1) and seg.right_pad_dims_to(t, t) is t
    assert seg.with_decode_head and not seg.with_auxiliary_head and not seg.with_neck
    dep = _depth_model()
    dp = dep._get_sampling_timesteps(2, device="cpu")
    assert len(dp) == 3 and torch.equal(dp[0], torch.tensor([[1.0, 1.0], [1 - 2 / 3, 1 - 2 / 3]]))
    assert float(dp[-1][1, 0]) == 0.0
    tt = torch.tensor([0.0, 0.5, 1.0])
    assert torch.equal(dep.gamma(tt), torch.cos(((tt + 0.0002) / 1.00025) * torch.pi / 2) ** 2)


def test_constructor_errors_mirror_reference():
    cfg = Config.fromfile(os.path.join(HERE, "fixtures", "ddp_toy_config.py"))
    bad = dict(cfg.model)
    bad["noise_schedule"] = "quadratic"
    with pytest.raises(ValueError, match="invalid noise schedule"):       # ddp.py:90
        build_segmentor(bad)
    bad = dict(cfg.model)
    bad["decode_head"] = dict(cfg.model.decode_head, positional_encoding=dict(type="SinePositionalEncoding", num_feats=64,
                                                                               normalize=True, offset=-0.5))
    with pytest.raises(AssertionError, match="embed_dims should be exactly 2 times"):   # deformable_head_with_time.py:45-47
        build_segmentor(bad)
    with pytest.raises(KeyError):
        build_segmentor(dict(type="NoSuchSegmentor"))


def test_neck_plugin_surface():
    """FPN / MultiStageMerging: reference constructor arguments and state-dict keys; arguments outside what the DDP
    configs use are refused loudly; without a GPU the forward fails loudly (no CPU fallback)."""
    gn = dict(type="GN", num_groups=32)
    fpn = MODELS.build(dict(type="FPN", in_channels=[96, 192, 384, 768], out_channels=256, act_cfg=None, norm_cfg=gn,
                            num_outs=4))
    msm = MODELS.build(dict(type="MultiStageMerging", in_channels=[256] * 4, out_channels=256, kernel_size=1,
                            norm_cfg=gn, act_cfg=None))
    assert isinstance(fpn, FPN) and isinstance(msm, MultiStageMerging)
    want = NO.make_weights([96, 192, 384, 768], seed=0)
    got = {"neck.0." + k: v for k, v in fpn.state_dict().items()}
    got.update({"neck.1." + k: v for k, v in msm.state_dict().items()})
    assert set(got) == set(want) and all(tuple(got[k].shape) == tuple(want[k].shape) for k in want)
    for bad in (dict(num_outs=5), dict(add_extra_convs=True), dict(norm_cfg=dict(type="BN")), dict(norm_cfg=None),
                dict(act_cfg=dict(type="ReLU")), dict(upsample_cfg=dict(mode="bilinear")), dict(start_level=1)):
        with pytest.raises(NotImplementedError):
            MODELS.build({**dict(type="FPN", in_channels=[96, 192, 384, 768], out_channels=256, norm_cfg=gn, num_outs=4),
                          **bad})
    with pytest.raises(NotImplementedError):
        MODELS.build(dict(type="MultiStageMerging", in_channels=[256] * 4, out_channels=256, kernel_size=3, norm_cfg=gn))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            fpn([torch.zeros(1, c, 4, 4) for c in (96, 192, 384, 768)])


def _bev_head_cfg(scopes):
    return dict(type="DeformableHeadWithTime", in_channels=256, num_feature_levels=1,
                encoder=dict(type="DetrTransformerEncoder", num_layers=5, transformerlayers=dict(
                    type="BaseTransformerLayer", use_time_mlp=True,
                    attn_cfgs=dict(type="MultiScaleDeformableAttention", embed_dims=256, num_levels=1, num_heads=8, dropout=0.0),
                    ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024, num_fcs=2,
                                  act_cfg=dict(type="GELU"), ffn_drop=0.0),
                    operation_order=["self_attn", "norm", "ffn", "norm"])),
                positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5),
                classes=["drivable_area", "ped_crossing", "walkway", "stop_line", "carpark_area", "divider"], loss="focal",
                grid_transform=dict(input_scope=scopes[0], output_scope=scopes[1]))


def test_bev_plugin_surface():
    """BEV DDP fusion model + head: reference constructor arguments, state-dict keys and grid coordinates."""
    from ddp_b200.bev import FUSIONMODELS, BevDDP, grid_coords
    from oracle import bev_oracle as BO
    for feat in (256, 512):
        cfg = BO.BevConfig(feat_channels=feat)
        kw = dict(type="DDP", bit_scale=0.01, timesteps=3, randsteps=5, time_difference=1, learned_sinusoidal_dim=16,
                  sample_range=[0, 0.999], noise_schedule="cosine", diffusion="ddim",
                  encoders=dict(camera=None, lidar=None), fuser=None, decoder=dict(backbone=None, neck=None),
                  heads=dict(object=None, map=_bev_head_cfg((cfg.input_scope, cfg.output_scope))))
        if feat != 512:
            kw["feat_channels"] = feat
        model = FUSIONMODELS.build(kw)
        assert isinstance(model, BevDDP) and model.num_classes == 6 and model.threshold == 0.5
        want = BO.make_weights(cfg, seed=0)
        got = model.state_dict()
        assert set(got) == set(want) and all(tuple(got[k].shape) == tuple(want[k].shape) for k in want)
        gy, gx = model.heads["map"].grid_coords()
        oy, ox = BO.grid_coords(cfg.input_scope, cfg.output_scope)
        assert gy.numel() == 200 and torch.equal(gy, oy) and torch.equal(gx, ox)
    with pytest.raises(ValueError, match="invalid noise schedule"):                  # fusion_models/ddp.py:102
        FUSIONMODELS.build(dict(kw, noise_schedule="quadratic"))
    with pytest.raises(NotImplementedError):
        FUSIONMODELS.build(dict(kw, heads=dict(object=dict(type="TransFusionHead"), map=None)))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            model.ddim_sample([torch.zeros(1, 512, 128, 128)])
    assert all(torch.equal(a, b) for a, b in zip(grid_coords(cfg.input_scope, cfg.output_scope), (oy, ox)))


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_reference_bev_yaml_configs_build():
    """The two shipped BEV DDP configs load UNCHANGED (torchpack-style recursive defaults + ${...} interpolation) and build."""
    from ddp_b200.bev import FUSIONMODELS
    from ddp_b200.config import load_yaml
    files = sorted(glob.glob(f"{REF}/bev/configs/nuscenes/seg/ddp-*.yaml"))
    assert len(files) == 2
    for f in files:
        cfg = load_yaml(f)
        m = cfg.model
        assert m.type == "DDP" and cfg.max_epochs == cfg.runner.max_epochs            # child overrides parent; ${max_epochs}
        assert m.heads.map.classes == cfg.map_classes and len(cfg.map_classes) == 6    # ${map_classes}
        assert m.heads.map.grid_transform.output_scope == [[-50, 50, 0.5], [-50, 50, 0.5]]      # from seg/default.yaml
        assert m.encoders.camera.vtransform.feature_size == [256 // 8, 704 // 8]      # ${[image_size[0] // 8, image_size[1] // 8]}
        model = FUSIONMODELS.build(m)
        assert model.timesteps == m.timesteps and model.randsteps == m.randsteps and model.bit_scale == 0.01
        assert model.feat_channels == m.get("feat_channels", 512)
        assert model.heads["map"].encoder.num_layers == 5
        assert model.heads["map"].grid_coords()[0].numel() == 200


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_reference_config_files_build_unchanged():
    files = sorted(glob.glob(f"{REF}/segmentation/configs/*/ddp_*.py"))
    assert len(files) == 14
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in files:
            cfg = Config.fromfile(f)
            model = build_segmentor(cfg.model)
            assert type(model).__name__ == cfg.model.type
            assert model.timesteps == cfg.model.timesteps and model.bit_scale == 0.01
            # the neck pair of every config builds into the fused CUDA neck with the reference's state-dict keys
            assert isinstance(model.neck, FusedNeck)
            want = NO.make_weights(list(cfg.model.neck[0].in_channels), seed=0)
            got = {k: v for k, v in model.state_dict().items() if k.startswith("neck.")}
            assert set(got) == set(want) and all(tuple(got[k].shape) == tuple(want[k].shape) for k in want)
        dfiles = sorted(glob.glob(f"{REF}/depth/configs/ddp_*/*.py"))
        assert len(dfiles) == 8
        for f in dfiles:
            cfg = Config.fromfile(f)
            model = build_depther(cfg.model)
            assert model.max_depth == cfg.model.max_depth
            assert isinstance(model.neck, FusedNeck)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_drop_in_inside_the_reference_tree():
    """INTEGRATION.md scenario A on the host side, inside the reference's own mmseg (subprocess: the import shim
    rewires `mmcv`): register_into_mmseg() -> the reference's build_segmentor builds the unchanged config into
    ddp_b200.DDP with the reference's own Swin backbone and the fused CUDA neck; a reference state dict loads by key."""
    import subprocess
    import sys
    res = subprocess.run([sys.executable, os.path.join(HERE, "fixtures", "inside_reference.py")], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "INSIDE-REFERENCE-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
    # the depth toolbox: reference build_depther + unchanged NYU config + the reference's own Swin backbone
    res = subprocess.run([sys.executable, os.path.join(HERE, "fixtures", "inside_reference_depth.py")], capture_output=True,
                         text=True, timeout=600)
    assert res.returncode == 0 and "INSIDE-REFERENCE-DEPTH-OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_shard_bounds_cover_batch():
    for B in (1, 5, 8, 64):
        for W in (1, 2, 3, 8):
            spans = [D.shard_bounds(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_bind_near_gpu_intersects_nvml_mask_and_never_raises(monkeypatch):
    """bench.py binds each rank to its GPU's NUMA node when N > 1: the helper must take the NVML mask, intersect it with
    what the process may use, do nothing when that changes nothing, and swallow every failure (no NVML here)."""
    import sys
    import types
    assert D.bind_near_gpu(0) is None                      # this container: no NVML / no GPU
    have = sorted(os.sched_getaffinity(0))
    calls = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append((pid, set(cpus))))
    fake = types.ModuleType("pynvml")
    fake.nvmlInit = lambda: None
    fake.nvmlDeviceGetHandleByUUID = lambda u: (_ for _ in ()).throw(RuntimeError("no uuid"))
    fake.nvmlDeviceGetHandleByIndex = lambda i: i
    mask = {"words": None}
    fake.nvmlDeviceGetCpuAffinity = lambda h, n: mask["words"]
    monkeypatch.setitem(sys.modules, "pynvml", fake)
    words = [0] * ((max(have) // 64) + 1)
    for c in have:
        words[c // 64] |= 1 << (c % 64)
    mask["words"] = list(words)
    assert D.bind_near_gpu(0) is None and not calls       # the mask is everything we already have: nothing to do
    if len(have) > 1:
        words[have[-1] // 64] &= ~(1 << (have[-1] % 64))   # GPU-local set = all but the last CPU
        mask["words"] = list(words)
        assert D.bind_near_gpu(0) == have[:-1]
        assert calls == [(0, set(have[:-1]))]
    mask["words"] = [0] * len(words)                       # empty intersection: leave the process alone
    calls.clear()
    assert D.bind_near_gpu(0) is None and not calls


def _gloo_worker(rank, world, port, batch, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        x = torch.randn(batch, 4, 3, 5, generator=g)
        noise = torch.randn(batch, 2, 4, 3, 5, generator=g)

        def fake_sample(xl, nl):                 # stands in for engine.sample on this rank's shard
            return xl * 2 + nl.mean(1)
        out = D.distributed_sample(fake_sample, x, noise)
        ok = torch.equal(out, x * 2 + noise.mean(1))
        # the small-payload form: a per-rank post step (stands in for engine.resize_argmax) and ONE gather of uint8 class maps
        cls = D.distributed_sample(fake_sample, x, noise, post=lambda lg: lg.argmax(1).to(torch.uint8))
        ok = ok and cls.dtype == torch.uint8 and torch.equal(cls, (x * 2 + noise.mean(1)).argmax(1).to(torch.uint8))
        q.put((rank, ok, tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5])
def test_two_rank_gloo_shard_and_gather(batch):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29511 + batch
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(shape == (batch, 4, 3, 5) for _, _, shape in res)
