"""GPU parity tests: the CUDA path (through the C ABI) against (1) the golden outputs of the
unmodified reference and (2) the oracle, at sizes the oracle finishes in seconds, plus
size-independent properties at larger sizes."""
import os

import numpy as np
import pytest
import torch

from oracle import ddp_oracle as O
from golden_util import golden_files, load_case
import parity as P

pytestmark = pytest.mark.gpu

MODES = ["fp32", "tc_3xf16"]


def make_engine(cfg: O.OracleConfig, W, mode="fp32"):
    from ddp_b200 import DecodeEngine
    eng = DecodeEngine(task=cfg.task, num_classes=cfg.num_classes, timesteps=cfg.timesteps,
                       time_difference=cfg.time_difference, sample_range=cfg.sample_range,
                       noise_schedule=cfg.noise_schedule, accumulation=cfg.accumulation, diffusion=cfg.diffusion,
                       bit_scale=cfg.bit_scale, num_layers=cfg.num_layers, min_depth=cfg.min_depth,
                       max_depth=cfg.max_depth, gemm_mode=mode)
    eng.load_state_dict(W)
    return eng


def argmax_report(a, b):
    """-> (#mismatching pixels, #pixels)."""
    am, bm = a.argmax(1), b.argmax(1)
    return int((am != bm).sum()), am.numel()


# tolerance on fp32 results computed in a different summation order (values are O(1)); tests/parity.py holds the rule
ATOL = P.ATOL
# Teacher-forced / single-evaluation comparisons (no feedback): a class-map pixel may differ from the oracle only if the
# ORACLE's own top-2 logit margin there is below TIE_TOL = 2 x the fp32 oracle's distance from its own fp64 evaluation
# (5.9e-5).  Whole-loop comparisons go through parity.check_seg_parity (fp64-adjudicated, closed loop).
TIE_TOL = 1.2e-4


def check_class_map(got_logits, ref_logits, what, dim=1, tie_tol=TIE_TOL):
    """argmax over `dim` must agree except at reference ties; returns the number of tie flips."""
    ga, ra = got_logits.argmax(dim), ref_logits.argmax(dim)
    bad = ga != ra
    n_bad = int(bad.sum())
    if n_bad:
        top2 = ref_logits.topk(2, dim=dim).values
        margin = (top2.select(dim, 0) - top2.select(dim, 1))[bad]
        P.log_record(dict(what=what, rule="single_evaluation_tie", flips=n_bad, pixels=bad.numel(), max_flip_margin=float(margin.max())))
        assert float(margin.max()) < tie_tol, f"{what}: {n_bad} class-map pixels differ, margin up to {float(margin.max()):.3e}"
    return n_bad


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("seg"), ids=lambda p: os.path.basename(p)[:-4])
def test_seg_matches_reference_golden(path, mode):
    cfg, W, x, noise, g = load_case(path)
    eng = make_engine(cfg, W, mode)
    eng.plan(1, cfg.randsteps, x.shape[2], x.shape[3])
    taps = [eng.add_tap(6, k, -1, cfg.num_classes) for k in range(cfg.timesteps)]   # DDP_TAP_LOGITS
    dn = torch.as_tensor(g["ddpm_noise"]) if cfg.diffusion == "ddpm" else None        # (1,T,R,256,h,w)
    sn = dn[0][:, None].cuda() if dn is not None else None                            # (T,1,R,256,h,w)
    out, cls = eng.sample(x.cuda(), noise.cuda(), return_cls=True, step_noise=sn)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["out"])
    out = out.cpu()
    R, h, w, C = cfg.randsteps, x.shape[2], x.shape[3], cfg.num_classes
    # per-step logits and class maps (the index work of the loop) against the REFERENCE's own per-step logits
    steps_exact = True
    for k in range(cfg.timesteps):
        lg = taps[k].cpu().view(R, h * w, C)
        ref_lg = torch.from_numpy(g["step_logits"][k]).permute(0, 2, 3, 1).reshape(R, h * w, C)
        steps_exact = steps_exact and P.class_maps_equal(lg, ref_lg, dim=2) and (lg - ref_lg).abs().max().item() < ATOL
    what = f"{os.path.basename(path)} [{mode}]"
    # every step against the oracle (pinned bit-for-bit to these goldens) on the CUDA path's own input, and the final output
    # against the golden
    P.check_seg_parity(eng, W, cfg, x, noise, what, ref=ref, ddpm_noise=dn, out=out)
    assert steps_exact, f"{what}: a per-step golden differed (the closed-loop rule adjudicated it: see the parity log)"
    assert torch.equal(cls.cpu().long(), out.argmax(1))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("path", golden_files("depth"), ids=lambda p: os.path.basename(p)[:-4])
def test_depth_matches_reference_golden(path, mode):
    cfg, W, x, noise, g = load_case(path)
    eng = make_engine(cfg, W, mode)
    out = eng.sample(x.cuda(), noise.cuda()).cpu()
    ref = torch.from_numpy(g["out"])
    # north-star tolerance for depth: |delta| < 1e-3 (metres, after the clamp)
    assert (out - ref).abs().max().item() < 1e-3
    assert (out - ref).abs().max().item() < ATOL       # and in fact at fp32 rounding level


@pytest.mark.parametrize("mode", MODES)
def test_every_layer_against_oracle_teacher_forced(mode):
    """Per-layer intermediates of every step, each step started from the ORACLE's state, so kernel
    error is separated from the feedback cascade."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3, randsteps=2)
    W = O.make_weights(cfg, seed=31)
    B, h, w = 2, 10, 14
    x, noise = O.make_inputs(cfg, B, h, w, seed=77)
    ref, traces = O.sample(W, cfg, x, noise, trace=True)
    eng = make_engine(cfg, W, mode)
    eng.plan(B, cfg.randsteps, h, w)
    R, N = cfg.randsteps, h * w
    for k in range(1, cfg.timesteps):      # step k starts from the oracle's state after step k-1
        st = torch.stack([traces[b].mask_t[k - 1] for b in range(B)])      # (B,R,256,h,w)
        eng.set_state_override(k, st.cuda())
    kinds = {"value": (1, 256), "sampling": (2, 96), "gathered": (3, 256), "ln1": (4, 256), "out": (5, 256)}
    bufs = {}
    for k in range(cfg.timesteps):
        bufs[("head_in", k)] = eng.add_tap(0, k, -1, 256)
        bufs[("logits", k)] = eng.add_tap(6, k, -1, cfg.num_classes)
        bufs[("state", k)] = eng.add_tap(7, k, -1, 256)
        for j in range(cfg.num_layers):
            for name, (kind, width) in kinds.items():
                bufs[(name, k, j)] = eng.add_tap(kind, k, j, width)
    out = eng.sample(x.cuda(), noise.cuda()).cpu()
    worst = {}

    def cmp(key, got, want, tol):
        d = (got - want).abs().max().item()
        worst[key[0]] = max(worst.get(key[0], 0.0), d)
        assert d < tol, f"{key}: max |d| = {d:.3e}"

    def tok(t):    # (R,C,h,w) -> (R,N,C)
        return t.flatten(2).transpose(1, 2)

    for k in range(cfg.timesteps):
        for b in range(B):
            tr = traces[b]
            sl = slice(b * R, (b + 1) * R)
            cmp(("head_in", k), bufs[("head_in", k)].cpu().view(B * R, N, 256)[sl], tok(tr.feat[k]), 1e-4)
            for j in range(cfg.num_layers):
                t = tr.layers[k][j]
                samp = torch.cat([t["offsets"], t["attn"]], dim=2)
                cmp(("value", k, j), bufs[("value", k, j)].cpu().view(B * R, N, 256)[sl], t["value"], 1e-4)
                cmp(("sampling", k, j), bufs[("sampling", k, j)].cpu().view(B * R, N, 96)[sl], samp, 1e-4)
                cmp(("gathered", k, j), bufs[("gathered", k, j)].cpu().view(B * R, N, 256)[sl], t["gathered"], 1e-4)
                cmp(("ln1", k, j), bufs[("ln1", k, j)].cpu().view(B * R, N, 256)[sl], t["ln1"], 1e-4)
                cmp(("out", k, j), bufs[("out", k, j)].cpu().view(B * R, N, 256)[sl], t["out"], 2e-4)
            lg = bufs[("logits", k)].cpu().view(B * R, N, cfg.num_classes)[sl]
            cmp(("logits", k), lg, tok(tr.logits[k]), 2e-4)
            check_class_map(lg, tok(tr.logits[k]), f"teacher-forced step {k} image {b} [{mode}]", dim=2)
            same = (lg.argmax(2) == tok(tr.logits[k]).argmax(2))[..., None]          # the DDIM update follows the class map
            cmp(("state", k), bufs[("state", k)].cpu().view(B * R, N, 256)[sl] * same, tok(tr.mask_t[k]) * same, 1e-5)
    # every step was started from the oracle's state, so the final map is ONE evaluation on the oracle's input
    assert (out - ref).abs().max().item() < ATOL
    check_class_map(out, ref, f"teacher-forced final [{mode}]")
    print("worst |d| per tensor:", {k: f"{v:.2e}" for k, v in worst.items()})


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("case", [
    dict(task="seg", num_classes=150, T=3, R=1, acc=True, B=2, h=16, w=16),
    dict(task="seg", num_classes=19, T=10, R=1, acc=False, B=2, h=8, w=24),
    dict(task="seg", num_classes=19, T=2, R=8, acc=False, B=1, h=9, w=9),      # uncertainty mode, K=8 samples
    dict(task="seg", num_classes=1, T=2, R=1, acc=False, B=1, h=5, w=6),       # single class
    dict(task="seg", num_classes=256, T=1, R=1, acc=True, B=1, h=4, w=7),      # max classes
    dict(task="seg", num_classes=19, T=1, R=1, acc=False, B=1, h=1, w=1),      # one token
    dict(task="seg", num_classes=19, T=1, R=1, acc=False, B=1, h=1, w=67),     # one row, ragged
    dict(task="depth", num_classes=1, T=20, R=1, acc=False, B=2, h=12, w=16),
    dict(task="depth", num_classes=1, T=3, R=3, acc=False, B=2, h=7, w=5),
], ids=lambda c: f"{c['task']}_C{c['num_classes']}_T{c['T']}_R{c['R']}_B{c['B']}_{c['h']}x{c['w']}")
def test_against_oracle_end_to_end(case, mode):
    cfg = O.OracleConfig(task=case["task"], num_classes=case["num_classes"], timesteps=case["T"],
                         randsteps=case["R"], accumulation=case["acc"],
                         bit_scale=0.01 if case["task"] == "seg" else 0.1)
    W = O.make_weights(cfg, seed=40 + case["T"])
    x, noise = O.make_inputs(cfg, case["B"], case["h"], case["w"], seed=500 + case["R"])
    ref = O.sample(W, cfg, x, noise)
    eng = make_engine(cfg, W, mode)
    out = eng.sample(x.cuda(), noise.cuda()).cpu()
    assert out.shape == ref.shape
    d = (out - ref).abs().max().item()
    if cfg.task == "seg":
        P.check_seg_parity(eng, W, cfg, x, noise, f"oracle e2e {case} [{mode}]", ref=ref, out=out)
    else:
        assert d < 1e-3 and d < ATOL, f"max |d| = {d:.3e}"


def test_host_buffer_entry_point_equals_device_entry_point():
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=3)
    x, noise = O.make_inputs(cfg, 2, 8, 8, seed=9)
    eng = make_engine(cfg, W)
    a = eng.sample(x.cuda(), noise.cuda()).cpu()
    cls = torch.empty((2, 8, 8), dtype=torch.int32).pin_memory()
    b = eng.sample_host(x.pin_memory(), noise.pin_memory(), cls=cls)
    assert torch.equal(a, b)
    assert torch.equal(cls.long(), a.argmax(1))


@pytest.mark.parametrize("task", ["seg", "depth"])
def test_host_pipeline_chunks_do_not_change_a_bit(task):
    """ddp_sample_host_ex pipelines the batch in groups of images against the PCIe copies: any chunk count (even / uneven
    groups, more chunks than images) must give the bits of the device-resident call; `out_device` receives the same."""
    cfg = O.OracleConfig(task=task, num_classes=19, timesteps=2, randsteps=2, bit_scale=0.01 if task == "seg" else 0.1)
    W = O.make_weights(cfg, seed=4)
    x, noise = O.make_inputs(cfg, 5, 9, 15, seed=12)
    eng = make_engine(cfg, W, "tc_3xf16")
    want, want_cls = (eng.sample(x.cuda(), noise.cuda(), return_cls=True) if task == "seg" else (eng.sample(x.cuda(), noise.cuda()), None))
    xh, nh = x.pin_memory(), noise.pin_memory()
    for chunks in (0, 1, 2, 3, 5, 9):
        cls = torch.zeros((5, 9, 15), dtype=torch.int32).pin_memory() if task == "seg" else None
        dev = torch.zeros_like(want)
        got = eng.sample_host(xh, nh, cls=cls, out_device=dev, chunks=chunks)
        assert torch.equal(got, want.cpu()), f"chunks={chunks}"
        assert torch.equal(dev, want)
        if task == "seg":
            assert torch.equal(cls, want_cls.cpu())
    # pageable host memory works too (the copies just do not overlap)
    assert torch.equal(eng.sample_host(x.clone(), noise.clone(), chunks=2), want.cpu())


def test_host_streaming_submit_wait_bitwise_and_ordering():
    """ddp_sample_host_submit / _wait: two calls in flight with different inputs give the bits of the synchronous entry
    point, a third submit without a wait is refused, tickets can be waited for out of order."""
    from ddp_b200._lib import DDPError
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=6)
    eng = make_engine(cfg, W, "tc_3xf16")
    ins = [tuple(t.pin_memory() for t in O.make_inputs(cfg, 3, 8, 12, seed=100 + i)) for i in range(5)]
    want = [eng.sample(x.cuda(), n.cuda()).cpu() for x, n in ins]
    outs = [torch.zeros_like(want[0]).pin_memory() for _ in ins]
    clss = [torch.zeros((3, 8, 12), dtype=torch.int32).pin_memory() for _ in ins]
    t0 = eng.submit_host(*ins[0], outs[0], cls=clss[0])
    t1 = eng.submit_host(*ins[1], outs[1], cls=clss[1])
    with pytest.raises(DDPError, match="already in flight"):
        eng.submit_host(*ins[2], outs[2])
    eng.wait_host(t1)                       # out of order
    eng.wait_host(t0)
    with pytest.raises(DDPError, match="not in flight"):
        eng.wait_host(t0)
    prev = None
    for i in range(2, 5):                   # steady state: submit i, then wait i-1
        t = eng.submit_host(*ins[i], outs[i], cls=clss[i])
        if prev is not None:
            eng.wait_host(prev)
        prev = t
    eng.wait_host(prev)
    for i in range(5):
        assert torch.equal(outs[i], want[i]), f"call {i}"
        assert torch.equal(clss[i].long(), want[i].argmax(1))
    # the synchronous entry point still works afterwards and agrees
    assert torch.equal(eng.sample_host(*ins[0]), want[0])


def test_batched_call_equals_per_image_calls_bitwise():
    """Images are independent: a batched call must equal per-image calls bit for bit,
    and stochastic samples r only meet in the final mean."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3, randsteps=2)
    W = O.make_weights(cfg, seed=8)
    x, noise = O.make_inputs(cfg, 4, 24, 40, seed=10)
    eng = make_engine(cfg, W)
    full = eng.sample(x.cuda(), noise.cuda()).cpu()
    for b in range(4):
        one = eng.sample(x[b:b + 1].cuda(), noise[b:b + 1].cuda()).cpu()
        assert torch.equal(full[b:b + 1], one)
    # swapping the two samples of an image leaves the mean unchanged up to one rounding
    sw = eng.sample(x.cuda(), noise.flip(1).cuda()).cpu()
    assert (sw - full).abs().max().item() < 1e-6


def test_full_size_properties_cityscapes_shape():
    """BASELINE config 3 shape (128x256 tokens, 19 classes): size-independent properties.
    accumulation output is a mean of softmaxes: rows sum to 1, values in [0,1]; determinism."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3, accumulation=True)
    W = O.make_weights(cfg, seed=5)
    x, noise = O.make_inputs(cfg, 2, 128, 256, seed=11)
    eng = make_engine(cfg, W)
    xc, nc = x.cuda(), noise.cuda()
    out, cls = eng.sample(xc, nc, return_cls=True)
    out2 = eng.sample(xc, nc)
    assert torch.equal(out, out2), "kernels must be deterministic"
    s = out.sum(1)
    assert (s - 1).abs().max().item() < 1e-5
    assert out.min().item() >= 0 and out.max().item() <= 1 + 1e-6
    assert torch.equal(cls.long(), out.argmax(1))
    # one image of the batch against the oracle would take ~1 min; a 1/16 crop of tokens cannot be
    # compared (attention is spatial), so check one image of the batch equals its solo run instead
    solo = eng.sample(xc[1:2], nc[1:2])
    assert torch.equal(solo, out[1:2])


def test_error_behaviour_mirrors_reference():
    from ddp_b200 import DecodeEngine
    from ddp_b200._lib import DDPError
    with pytest.raises(ValueError, match="invalid noise schedule"):        # ddp.py:90
        DecodeEngine(noise_schedule="quadratic")
    with pytest.raises(NotImplementedError):                               # ddp.py:123
        DecodeEngine(diffusion="euler")
    eng = DecodeEngine(num_classes=19)
    with pytest.raises(DDPError, match="ddp_commit_weights"):
        eng.plan(1, 1, 4, 4)
    with pytest.raises(KeyError):
        eng.load_state_dict({})
    cfg = O.OracleConfig(num_classes=19)
    W = O.make_weights(cfg, seed=1)
    bad = dict(W)
    bad["decode_head.conv_seg.weight"] = torch.zeros(20, 256, 1, 1)
    with pytest.raises(ValueError):
        eng.load_state_dict(bad)
    eng.load_state_dict({**W, "backbone.stem.weight": torch.zeros(3)})      # extra keys are ignored
    with pytest.raises(DDPError):
        DecodeEngine(num_classes=300)


def test_library_schedule_close_to_reference_schedule():
    """The C default schedule (plain float math) vs the host/torch one the plug-in hands over."""
    from ddp_b200 import DecodeEngine
    for T in (3, 10):
        a = DecodeEngine(num_classes=19, timesteps=T, host_schedule=False).get_schedule()
        b = DecodeEngine(num_classes=19, timesteps=T, host_schedule=True).get_schedule()
        for col_a, col_b in zip(a, b):
            assert np.allclose(col_a, col_b, rtol=2e-5, atol=2e-6), (col_a, col_b)


# ------------------------------------------------------------------------------------------------
# through the plug-in classes (the reference-facing surface)
# ------------------------------------------------------------------------------------------------
def _toy_model(timesteps=3, randsteps=1, **over):
    import warnings
    from ddp_b200.config import Config
    from ddp_b200.registry import build_segmentor
    import test_plugin_cpu  # noqa: F401  (registers ToyBackbone)
    cfg = Config.fromfile(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures", "ddp_toy_config.py"))
    m = dict(cfg.model)
    m.update(timesteps=timesteps, randsteps=randsteps, **over)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return build_segmentor(m)


@pytest.mark.parametrize("mode", MODES)
def test_plugin_ddim_sample_matches_oracle(mode):
    model = _toy_model(timesteps=3, randsteps=2, gemm_mode=mode)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3, randsteps=2)
    W = O.make_weights(cfg, seed=9)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    x, noise = O.make_inputs(cfg, 2, 12, 20, seed=21)
    out = model.ddim_sample(x.cuda(), None, noise=noise.cuda()).cpu()
    ref = O.sample(W, cfg, x, noise)
    P.check_seg_parity(model.engine(), W, cfg, x, noise, f"plugin ddim_sample [{mode}]", ref=ref, out=out)
    # noise drawn inside, like ddp.py:220
    torch.manual_seed(5)
    out2 = model.ddim_sample(x.cuda(), None)
    torch.manual_seed(5)
    drawn = torch.randn((2, 2, 256, 12, 20), device="cuda")
    ref2 = O.sample(W, cfg, x, drawn.cpu())
    P.check_seg_parity(model.engine(), W, cfg, x, drawn.cpu(), f"plugin ddim_sample, internal noise [{mode}]", ref=ref2, out=out2.cpu())


def test_plugin_engine_follows_parameter_changes_by_any_route():
    """The cached CUDA engine must be rebuilt when parameters change without going through load_state_dict: mmcv's
    load_checkpoint recurses over _load_from_state_dict, and in-place edits bypass every hook.  The
    (data_ptr, _version) fingerprint of ddp_b200/_stale.py catches both."""
    model = _toy_model(timesteps=2)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=81)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    x, noise = O.make_inputs(cfg, 1, 8, 10, seed=82)
    a = model.ddim_sample(x.cuda(), None, noise=noise.cuda())
    eng = model.engine()
    assert model.engine() is eng                                   # unchanged parameters: the engine is reused
    with torch.no_grad():                                          # in-place edit of one hot-path weight
        model.decode_head.conv_seg.weight.mul_(-1.0)
    b = model.ddim_sample(x.cuda(), None, noise=noise.cuda())
    assert model.engine() is not eng and not torch.equal(a, b)
    W2 = dict(W)
    W2["decode_head.conv_seg.weight"] = -W["decode_head.conv_seg.weight"]
    ref = O.sample(W2, cfg, x, noise)
    P.check_seg_parity(model.engine(), W2, cfg, x, noise, "engine after an in-place weight edit", ref=ref, out=b.cpu())
    # the route mmcv's load_checkpoint takes: per-module _load_from_state_dict, never Module.load_state_dict
    sd = {k: v.cuda() for k, v in W.items()}
    with torch.no_grad():
        for name, mod in model.named_modules():
            prefix = name + "." if name else ""
            mod._load_from_state_dict(sd, prefix, {}, False, [], [], [])
    c = model.ddim_sample(x.cuda(), None, noise=noise.cuda())
    assert torch.equal(c, a)


def test_plugin_simple_test_end_to_end():
    """img -> (toy) encoder -> ddim loop in the library -> x4 bilinear resize -> softmax -> argmax -> numpy,
    i.e. BaseSegmentor.forward(return_loss=False) as tools/test.py calls it."""
    model = _toy_model(timesteps=2)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=12)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    img = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(2)).cuda()
    metas = [dict(img_shape=(64, 96, 3), ori_shape=(64, 96, 3), pad_shape=(64, 96, 3), flip=False)]
    torch.manual_seed(77)
    pred = model(img=[img], img_metas=[metas], return_loss=False)
    assert isinstance(pred, list) and pred[0].shape == (64, 96)
    with torch.no_grad():
        x = model.extract_feat(img)[0]
    torch.manual_seed(77)
    noise = torch.randn((1, 1, 256, 16, 24), device="cuda")
    logits = O.sample(W, cfg, x.cpu(), noise.cpu())
    up = torch.nn.functional.interpolate(logits, size=(64, 96), mode="bilinear", align_corners=False)
    want = up.softmax(1).argmax(1)[0].numpy()
    assert (pred[0] != want).mean() < 1e-3, f"{(pred[0] != want).sum()} of {want.size} pixels differ"


def test_plugin_depth_sample_matches_oracle():
    from ddp_b200.registry import build_depther
    import test_plugin_cpu  # noqa: F401
    dh = dict(type="DeformableHeadWithTime", in_channels=[256], channels=256, in_index=[0], dropout_ratio=0.,
              min_depth=1e-3, max_depth=10, num_feature_levels=1,
              encoder=dict(type="DetrTransformerEncoder", num_layers=6, transformerlayers=dict(
                  type="BaseTransformerLayer", use_time_mlp=True,
                  attn_cfgs=dict(type="MultiScaleDeformableAttention", embed_dims=256, num_levels=1, num_heads=8, dropout=0.),
                  ffn_cfgs=dict(type="FFN", embed_dims=256, feedforward_channels=1024, ffn_drop=0., act_cfg=dict(type="GELU")),
                  operation_order=("self_attn", "norm", "ffn", "norm"))),
              positional_encoding=dict(type="SinePositionalEncoding", num_feats=128, normalize=True, offset=-0.5))
    model = build_depther(dict(type="DDP", bit_scale=0.1, timesteps=4, min_depth=1e-3, max_depth=10,
                               backbone=dict(type="ToyBackbone"), decode_head=dh))
    cfg = O.OracleConfig(task="depth", timesteps=4, bit_scale=0.1)
    W = O.make_weights(cfg, seed=14)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    x, noise = O.make_inputs(cfg, 2, 10, 12, seed=22)
    out = model.sample(x.cuda(), None, noise=noise.cuda()).cpu()
    ref = O.sample(W, cfg, x, noise)
    assert (out - ref).abs().max().item() < 1e-3


def test_fast_mode_reports_mismatch_rate():
    """DDP_GEMM_TC_F16 (one fp16 MMA per product) is NOT parity-grade: it must stay close, and we record how close."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3)
    W = O.make_weights(cfg, seed=31)
    x, noise = O.make_inputs(cfg, 2, 32, 48, seed=78)
    ref = O.sample(W, cfg, x, noise)
    out = make_engine(cfg, W, "tc_f16").sample(x.cuda(), noise.cuda()).cpu()
    bad, tot = argmax_report(out, ref)
    d = (out - ref).abs().max().item()
    print(f"tc_f16: max|d|={d:.3e}, class-map mismatches {bad}/{tot}")
    assert d < 0.1 and bad / tot < 0.02


@pytest.mark.parametrize("mode", MODES)
def test_decode_head_forward_single_call(mode):
    """decode_head.forward(inputs, times) (deformable_head_with_time.py:90-132): one denoiser evaluation for a given
    time embedding, against the oracle's head_seg / head_depth."""
    model = _toy_model(timesteps=3, gemm_mode=mode)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=3)
    W = O.make_weights(cfg, seed=17)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(3, 256, 9, 13, generator=g)
    temb = torch.randn(1, 1024, generator=g)
    with torch.no_grad():
        want = O.head_seg(W, cfg, feat, temb)
    got = model.decode_head([feat.cuda()], temb.cuda()).cpu()
    assert got.shape == want.shape
    assert (got - want).abs().max().item() < ATOL
    check_class_map(got, want, f"decode_head.forward [{mode}]")


def test_plugin_ddpm_sample_matches_oracle():
    """diffusion='ddpm' (ddp.py:248-290): the plug-in draws the initial and the per-step noise in the reference's order."""
    model = _toy_model(timesteps=4, randsteps=2, diffusion="ddpm")
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=4, randsteps=2, diffusion="ddpm")
    W = O.make_weights(cfg, seed=19)
    model.load_state_dict(W, strict=False)
    model = model.cuda().eval()
    x, _ = O.make_inputs(cfg, 2, 9, 11, seed=23)
    torch.manual_seed(11)
    out = model.ddpm_sample(x.cuda(), None).cpu()
    torch.manual_seed(11)
    noise = torch.randn((2, 2, 256, 9, 11), device="cuda")
    steps = torch.stack([torch.randn_like(noise) for _ in range(4)])          # (T,B,R,256,h,w)
    dn = steps.permute(1, 0, 2, 3, 4, 5).cpu()
    ref = O.sample(W, cfg, x, noise.cpu(), ddpm_noise=dn)
    P.check_seg_parity(model.engine(), W, cfg, x, noise.cpu(), "plugin ddpm_sample", ref=ref, ddpm_noise=dn, out=out)


def test_fused_post_loop_tail_matches_torch():
    """ddp_resize_argmax == argmax(softmax(F.interpolate(logits, bilinear, align_corners=False))) (encoder_decoder.py:229-304)."""
    from ddp_b200 import DecodeEngine
    eng = DecodeEngine(num_classes=19)
    g = torch.Generator().manual_seed(1)
    for (B, C, h, w, H, W) in [(2, 19, 32, 64, 128, 256), (1, 150, 16, 16, 64, 64), (1, 19, 7, 13, 30, 50), (1, 3, 5, 5, 5, 5)]:
        logits = torch.randn(B, C, h, w, generator=g).cuda()
        got = eng.resize_argmax(logits, (H, W)).long()
        up = torch.nn.functional.interpolate(logits, size=(H, W), mode="bilinear", align_corners=False)
        want = up.softmax(1).argmax(1)
        diff = got != want
        if diff.any():          # only exact-rounding ties may differ
            top2 = up.topk(2, dim=1).values
            assert float((top2[:, 0] - top2[:, 1])[diff].max()) < 1e-5
        assert diff.float().mean().item() < 1e-4


def test_fused_inference_tail_flip_rescale_and_aug_test():
    """ddp_tail_probs / ddp_probs_argmax behind `inference` and `aug_test` (encoder_decoder.py:229-304): resize to the
    network input size, crop to img_shape, resize to ori_shape, softmax, flip back, sum over augmented views — against the
    eager torch path of the same plug-in on the same logits (same seed => same noise)."""
    model = _toy_model(timesteps=2)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    model.load_state_dict(O.make_weights(cfg, seed=71), strict=False)
    model = model.cuda().eval()
    g = torch.Generator().manual_seed(5)

    def both(fn):
        model.fused_tail = True
        torch.manual_seed(9)
        a = fn()
        model.fused_tail = False
        torch.manual_seed(9)
        b = fn()
        model.fused_tail = True
        return a, b

    views = [((64, 96), (60, 90), False, None), ((96, 128), (90, 128), True, "horizontal"), ((48, 64), (48, 61), True, "vertical")]
    ori = (75, 110, 3)
    imgs, metas = [], []
    for (ih, iw), (ch, cw), flip, d in views:
        imgs.append(torch.randn(1, 3, ih, iw, generator=g).cuda())
        metas.append([dict(img_shape=(ch, cw, 3), ori_shape=ori, pad_shape=(ih, iw, 3), flip=flip, flip_direction=d)])
    for img, meta in zip(imgs, metas):
        for rescale in (True, False):
            a, b = both(lambda: model.inference(img, meta, rescale))
            assert a.shape == b.shape
            assert (a - b).abs().max().item() < 2e-6, (meta, rescale, (a - b).abs().max().item())
            assert (a.sum(1) - 1).abs().max().item() < 1e-5
    a, b = both(lambda: model.aug_test(imgs, metas, rescale=True))
    assert a[0].shape == (75, 110) and b[0].shape == (75, 110)
    diff = torch.from_numpy(a[0] != b[0])
    if diff.any():          # only ties of the summed probabilities (inside their 1e-6 rounding) may differ
        model.fused_tail = False
        torch.manual_seed(9)
        psum = sum(model.inference(i, m, True) for i, m in zip(imgs, metas))[0].cpu()
        model.fused_tail = True
        top2 = psum.topk(2, dim=0).values
        assert float((top2[0] - top2[1])[diff].max()) < 1e-5


def test_unfused_ffn_pair_still_correct(monkeypatch):
    """DDP_B200_FUSE_FFN=0 selects the separate FFN1 (16-warp GELU epilogue) and FFN2 (LN + FiLM epilogue) kernels."""
    monkeypatch.setenv("DDP_B200_FUSE_FFN", "0")
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=33)
    x, noise = O.make_inputs(cfg, 2, 12, 16, seed=80)
    ref = O.sample(W, cfg, x, noise)
    eng = make_engine(cfg, W, "tc_3xf16")
    eng.profile(True)
    out = eng.sample(x.cuda(), noise.cuda()).cpu()
    prof = eng.profile_collect()
    assert prof["ffn1_gelu"][1] == 12 and prof["ffn2_ln_film"][1] == 12 and prof["ffn_fused"][1] == 0
    eng.profile(False)
    P.check_seg_parity(eng, W, cfg, x, noise, "unfused FFN pair", ref=ref, out=out)


@pytest.mark.parametrize("env", [{"DDP_B200_FFN_PAIR": "0"}, {"DDP_B200_GEMM_PAIR": "7"},
                                 {"DDP_B200_FFN_PAIR": "0", "DDP_B200_GEMM_PAIR": "0"}, {"DDP_B200_QPROJ_FUSED": "1"},
                                 {"DDP_B200_QPROJ_FUSED": "0"}],
                         ids=["ffn_single_cta", "all_gemm_pairs", "no_pairs", "qproj_fused", "qproj_separate"])
def test_cta_pair_options_agree_with_oracle(monkeypatch, env):
    """The fused FFN and the output projection run on CTA pairs (cta_group::2) by default, value / sampling on single
    CTAs; the other combinations (DDP_B200_FFN_PAIR=0, DDP_B200_GEMM_PAIR bit mask) must give the same answer.  The ragged grid
    (9 x 15 tokens, 2 images: 270 rows = 3 row tiles) leaves the second CTA of the last pair without rows."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=35)
    x, noise = O.make_inputs(cfg, 2, 9, 15, seed=82)
    ref = O.sample(W, cfg, x, noise)
    eng = make_engine(cfg, W, "tc_3xf16")
    out = eng.sample(x.cuda(), noise.cuda()).cpu()
    P.check_seg_parity(eng, W, cfg, x, noise, f"pair options {env}", ref=ref, out=out)


def test_full_size_single_step_against_oracle():
    """BASELINE config-3 token grid (128 x 256 = 32768 tokens, 19 classes), one image, one DDIM step: the whole
    per-step path at full size against the oracle (about 10-20 s of CPU work)."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=1)
    W = O.make_weights(cfg, seed=41)
    x, noise = O.make_inputs(cfg, 1, 128, 256, seed=90)
    ref = O.sample(W, cfg, x, noise)
    out = make_engine(cfg, W, "tc_3xf16").sample(x.cuda(), noise.cuda()).cpu()
    d = (out - ref).abs()
    print(f"full-size step: max|d|={d.max().item():.2e} mean|d|={d.mean().item():.2e}")
    assert d.max().item() < ATOL
    check_class_map(out, ref, "full-size 128x256 single step [tc_3xf16]")


def test_full_size_depth_properties():
    """NYU shape (120 x 160 tokens), T=20: output inside [min_depth, max_depth], deterministic, batch = per-image."""
    cfg = O.OracleConfig(task="depth", timesteps=20, bit_scale=0.1)
    W = O.make_weights(cfg, seed=42)
    x, noise = O.make_inputs(cfg, 2, 120, 160, seed=91)
    eng = make_engine(cfg, W, "tc_3xf16")
    a = eng.sample(x.cuda(), noise.cuda())
    b = eng.sample(x.cuda(), noise.cuda())
    assert torch.equal(a, b)
    assert a.min().item() >= cfg.min_depth - 1e-7 and a.max().item() <= cfg.max_depth + 1e-6
    solo = eng.sample(x[1:2].cuda(), noise[1:2].cuda())
    assert torch.equal(solo, a[1:2])


# ------------------------------------------------------------------------------------------------
# BASELINE.json's shapes at full size, whole loop, against the oracle (VERDICT r1 "weak" 3): the feedback over 10-20 steps
# at 16-32 k tokens is where near-tie flips would accumulate.  One image each (the oracle needs ~1 s per step and image
# on the GPU box's 16 host cores); batched = per-image is a separate bitwise test above.
# ------------------------------------------------------------------------------------------------
def test_full_size_cfg3_cityscapes_T10_against_oracle():
    """BASELINE config 3: 128 x 256 tokens, 19 classes, T = 10 — the headline loop, all ten steps, fp64-adjudicated."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=10)
    W = O.make_weights(cfg, seed=43)
    x, noise = O.make_inputs(cfg, 1, 128, 256, seed=92)
    ref = O.sample(W, cfg, x, noise)           # open loop as well: the record then says how the FINAL map compares
    P.check_seg_parity(make_engine(cfg, W, "tc_3xf16"), W, cfg, x, noise, "BASELINE cfg3 128x256 C19 T10 [tc_3xf16]", ref=ref)


def test_full_size_cfg2_ade_T3_accumulation_against_oracle():
    """BASELINE config 2: 128 x 128 tokens, 150 classes, T = 3 with accumulation (mean of softmaxes)."""
    cfg = O.OracleConfig(task="seg", num_classes=150, timesteps=3, accumulation=True)
    W = O.make_weights(cfg, seed=44)
    x, noise = O.make_inputs(cfg, 1, 128, 128, seed=93)
    ref = O.sample(W, cfg, x, noise)
    P.check_seg_parity(make_engine(cfg, W, "tc_3xf16"), W, cfg, x, noise, "BASELINE cfg2 128x128 C150 T3 acc [tc_3xf16]", ref=ref)


def test_full_size_cfg4_depth_T20_against_oracle():
    """BASELINE config 4: NYU 120 x 160 tokens, T = 20: |delta| < 1e-3 m (north-star tolerance) after 20 feedback steps."""
    cfg = O.OracleConfig(task="depth", timesteps=20, bit_scale=0.1)
    W = O.make_weights(cfg, seed=45)
    x, noise = O.make_inputs(cfg, 1, 120, 160, seed=94)
    ref = O.sample(W, cfg, x, noise)
    out = make_engine(cfg, W, "tc_3xf16").sample(x.cuda(), noise.cuda()).cpu()
    d = (out - ref).abs()
    P.log_record(dict(what="BASELINE cfg4 depth 120x160 T20 [tc_3xf16]", rule="depth |d| < 1e-3", max_abs_d_out=float(d.max()),
                      mean_abs_d_out=float(d.mean()), pixels=d.numel()))
    assert d.max().item() < 1e-3, f"max |d| = {d.max().item():.3e} m"


def test_full_size_cfg5_uncertainty_K8_T10_against_oracle():
    """BASELINE config 5: K = 8 stochastic samples of one 128 x 256 image, T = 10; the mean over the 8 samples."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=10, randsteps=8)
    W = O.make_weights(cfg, seed=46)
    x, noise = O.make_inputs(cfg, 1, 128, 256, seed=95)
    P.check_seg_parity(make_engine(cfg, W, "tc_3xf16"), W, cfg, x, noise, "BASELINE cfg5 128x256 C19 T10 K8 [tc_3xf16]")


@pytest.mark.parametrize("case", [dict(task="seg", T=4, R=5, B=2, h=12, w=18, acc=False), dict(task="seg", T=3, R=3, B=1, h=9, w=9, acc=True),
                                  dict(task="depth", T=3, R=4, B=2, h=8, w=10, acc=False)],
                         ids=lambda c: f"{c['task']}_T{c['T']}_R{c['R']}{'_acc' if c['acc'] else ''}")
def test_uncertainty_maps_match_oracle_definition(case):
    """ddp_set_uncertainty_outputs: class-change counts over steps x samples and the last-step disagreement
    of the R samples (depth: their standard deviation), against oracle.uncertainty; asking for them does not change `out`."""
    cfg = O.OracleConfig(task=case["task"], num_classes=19, timesteps=case["T"], randsteps=case["R"], accumulation=case["acc"],
                         bit_scale=0.01 if case["task"] == "seg" else 0.1)
    W = O.make_weights(cfg, seed=51)
    x, noise = O.make_inputs(cfg, case["B"], case["h"], case["w"], seed=52)
    ref, changes, spread = O.uncertainty(W, cfg, x, noise)
    eng = make_engine(cfg, W, "tc_3xf16")
    plain = eng.sample(x.cuda(), noise.cuda())
    out, unc = eng.sample(x.cuda(), noise.cuda(), return_uncertainty=True)
    assert torch.equal(out, plain)
    if cfg.task == "seg":
        P.check_seg_parity(eng, W, cfg, x, noise, f"uncertainty {case}", ref=ref, out=out.cpu())
        assert int(changes.sum()) > 0, "the case must exercise class changes"
        # identical class maps at every step (checked above: exact) => identical counts
        assert torch.equal(unc["changes"].cpu(), changes)
        assert (unc["spread"].cpu() - spread).abs().max().item() < 1e-6
    else:
        assert (out.cpu() - ref).abs().max().item() < 1e-3
        assert (unc["spread"].cpu() - spread).abs().max().item() < 1e-4
        # the synthetic depth head is insensitive to the noise (spread ~1e-4 m), so also check the statistic itself on
        # the CUDA path's own last-step predictions
        B, R, N = case["B"], case["R"], case["h"] * case["w"]
        tap = eng.add_tap(6, cfg.timesteps - 1, -1, 1)
        _, unc2 = eng.sample(x.cuda(), noise.cuda(), return_uncertainty=True)
        own = tap.view(B, R, N).std(1, unbiased=False).view(B, case["h"], case["w"])
        eng.clear_debug()
        assert (unc2["spread"] - own).abs().max().item() < 1e-6


def test_replan_keeps_shape_constants_and_results():
    """A re-plan that only changes the batch / sample count reuses the shape-only constants (no realloc); changing the
    grid or the weights rebuilds them.  Results must be what a fresh engine gives."""
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2, randsteps=2)
    W = O.make_weights(cfg, seed=61)
    W2 = O.make_weights(cfg, seed=62)
    x, noise = O.make_inputs(cfg, 3, 8, 12, seed=63)
    x2, noise2 = O.make_inputs(cfg, 1, 6, 10, seed=64)
    eng = make_engine(cfg, W, "tc_3xf16")
    a3 = eng.sample(x.cuda(), noise.cuda())
    a1 = eng.sample(x[:1].cuda(), noise[:1].cuda())                      # batch 3 -> 1, same grid
    a1r = eng.sample(x[:1].cuda(), noise[:1, :1].cuda())                 # R 2 -> 1
    b = eng.sample(x2.cuda(), noise2.cuda())                             # other grid
    a3b = eng.sample(x.cuda(), noise.cuda())                             # back
    assert torch.equal(a3, a3b) and torch.equal(a1, a3[:1])
    fresh = make_engine(cfg, W, "tc_3xf16")
    assert torch.equal(fresh.sample(x2.cuda(), noise2.cuda()), b)
    assert torch.equal(fresh.sample(x[:1].cuda(), noise[:1, :1].cuda()), a1r)
    eng.load_state_dict(W2)                                              # new weights: PE * W must be rebuilt
    assert torch.equal(eng.sample(x.cuda(), noise.cuda()), make_engine(cfg, W2, "tc_3xf16").sample(x.cuda(), noise.cuda()))


def test_empty_batch():
    """Empty input: torch semantics at the Python layer (an empty result, nothing launched); the C ABI rejects B < 1."""
    from ddp_b200._lib import DDPError
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    eng = make_engine(cfg, O.make_weights(cfg, seed=2), "tc_3xf16")
    out, cls = eng.sample(torch.zeros(0, 256, 4, 4).cuda(), torch.zeros(0, 1, 256, 4, 4).cuda(), return_cls=True)
    assert tuple(out.shape) == (0, 19, 4, 4) and tuple(cls.shape) == (0, 4, 4)
    with pytest.raises(DDPError, match="must be >= 1"):
        eng.plan(0, 1, 4, 4)
