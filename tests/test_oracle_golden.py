"""Pin the oracle against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by tests/golden/make_golden.py from /root/reference's own classes + configs)."""
import os

import numpy as np
import pytest
import torch

from oracle import ddp_oracle as O
from oracle import neck_oracle as NO
from oracle import bev_oracle as BO
from golden_util import golden_files, load_case, load_neck_case, load_bev_case


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_reference_output(path):
    cfg, W, x, noise, g = load_case(path)
    ref = torch.from_numpy(g["out"])
    # plain tensors: equal up to fp32 rounding order (torch's matmul folds a non-contiguous 3-D
    # input into one mm only when an operand has requires_grad, which nn.Parameter weights do)
    dn = g.get("ddpm_noise")
    out_plain = O.sample(W, cfg, x, noise, ddpm_noise=dn)
    assert out_plain.shape == ref.shape
    assert (out_plain - ref).abs().max().item() < 2e-5
    # with the reference's parameter flags the op sequence is identical: bit-for-bit
    W = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    out, traces = O.sample(W, cfg, x, noise, trace=True, ddpm_noise=dn)
    # same ops in the same order on the same CPU: bit-for-bit
    assert torch.equal(out, ref), f"max |d| = {(out - ref).abs().max().item():.3e}"
    tr = traces[0]
    if cfg.task == "seg":
        steps = torch.stack(tr.logits)
        assert torch.equal(steps, torch.from_numpy(g["step_logits"]))
        am = torch.stack(tr.argmax).numpy().astype(np.int16)
        assert np.array_equal(am, g["step_argmax"])
    else:
        steps = torch.stack(tr.logits)
        assert torch.equal(steps, torch.from_numpy(g["step_depth"]))


def test_schedule_known_answers():
    """SURVEY.md section 8c known-answer table (reference formulas ddp.py:22-28, 204-213)."""
    cfg = O.OracleConfig(timesteps=3)
    pairs = O.time_pairs_seg(cfg)
    assert pairs[0] == (1.0, pytest.approx(1 / 3)) and pairs[1][1] == 0 and pairs[2][1] == 0
    t = torch.tensor([1.0, 2 / 3, 1 / 3, 0.0])
    l = O.log_snr_cosine(t)
    assert torch.allclose(l, torch.tensor([-18.9075, -1.0989, 1.0978, 11.5129]), atol=2e-3)
    a, s = O.alpha_sigma(l)
    assert torch.allclose(a, torch.tensor([7.8e-5, 0.499955, 0.865934, 0.999995]), atol=2e-5)
    assert torch.allclose(s, torch.tensor([1.0, 0.866052, 0.500159, 0.003162]), atol=2e-5)
    cfg10 = O.OracleConfig(timesteps=10)
    p10 = O.time_pairs_seg(cfg10)
    assert p10[0] == (1.0, pytest.approx(0.8)) and p10[-1] == (pytest.approx(0.1), 0)
    a10, _ = O.alpha_sigma(O.log_snr_cosine(torch.tensor([0.9, 0.5, 0.1])))
    assert torch.allclose(a10, torch.tensor([0.156473, 0.707024, 0.987645]), atol=2e-5)
    assert O.time_pairs_seg(O.OracleConfig(timesteps=1)) == [(1.0, 0)]
    g = O.gamma_depth(torch.tensor([1.0, 0.0]))
    assert g[0].item() == pytest.approx(6.146e-9, rel=2e-2) and g[1].item() == pytest.approx(0.999999881, abs=1e-7)
    d20 = O.time_pairs_depth(O.OracleConfig(task="depth", timesteps=20))
    assert d20[0] == (1.0, pytest.approx(0.9)) and d20[1] == (pytest.approx(0.95), pytest.approx(0.85))


def test_linear_schedule_known_answers():
    """noise_schedule='linear' (ddp.py:18-19, no shipped config uses it): values of the reference's own
    beta_linear_log_snr / log_snr_to_alpha_sigma, evaluated here through the import shim when this test was written."""
    t = torch.tensor([1.0, 2 / 3, 1 / 3, 0.0, 0.5, 0.9])
    want_l = torch.tensor([-10.000054359436035, -4.432733058929443, -0.7119865417480469, 9.21028995513916,
                           -2.4144582748413086, -8.099796295166016])
    want_a = torch.tensor([0.006737610790878534, 0.108362577855587, 0.5737246870994568, 0.9999499917030334,
                           0.2864904999732971, 0.017421504482626915])
    want_s = torch.tensor([0.9999772906303406, 0.9941114783287048, 0.8190482258796692, 0.009999752044677734,
                           0.9580830335617065, 0.9998482465744019])
    l = O.log_snr_linear(t)
    a, s = O.alpha_sigma(l)
    assert torch.allclose(l, want_l, rtol=2e-6, atol=0) and torch.allclose(a, want_a, rtol=2e-6, atol=0)
    assert torch.allclose(s, want_s, rtol=2e-6, atol=0)
    # the host schedule handed to the library takes the same branch
    from ddp_b200 import schedule as S
    ls, a3, s3, an3, sn3 = S.seg_schedule(3, 1, (0, 0.999), "linear")
    assert ls[0] == pytest.approx(float(want_l[0]), rel=2e-6) and a3[1] == pytest.approx(float(want_a[1]), rel=2e-6)
    assert an3[0] == pytest.approx(float(want_a[2]), rel=2e-6) and sn3[2] == pytest.approx(float(want_s[3]), rel=2e-6)


def test_batched_equals_per_image_loop():
    cfg = O.OracleConfig(task="seg", num_classes=7, timesteps=2, randsteps=2)
    W = O.make_weights(cfg, seed=3)
    x, noise = O.make_inputs(cfg, B=3, h=6, w=9, seed=5)
    full = O.sample(W, cfg, x, noise)
    for b in range(3):
        one = O.sample(W, cfg, x[b:b + 1], noise[b:b + 1])
        assert torch.equal(full[b:b + 1], one)


def test_msda_gather_is_index_space_bilinear():
    """The gather samples value at pixel (j+off_x, i+off_y), bilinear, zero padded
    (vmmcv/ops/multi_scale_deform_attn.py:94-151): check against a direct loop."""
    torch.manual_seed(0)
    h, w = 5, 7
    N = h * w
    value = torch.randn(1, N, O.HEADS, 32)
    off = torch.randn(1, N, O.HEADS, 1, O.POINTS, 2) * 3
    aw = torch.rand(1, N, O.HEADS, 1, O.POINTS)
    ref = O.reference_points(h, w, torch.float32)
    loc = ref[:, :, None, :, None, :] + off / torch.tensor([[w, h]])[None, None, None, :, None, :]
    got = O.msda_gather(value, h, w, loc, aw)
    want = torch.zeros(N, 256)
    v = value[0].reshape(h, w, O.HEADS, 32)
    for n in range(N):
        i, j = divmod(n, w)
        for m in range(O.HEADS):
            for p in range(O.POINTS):
                px = j + off[0, n, m, 0, p, 0].item()
                py = i + off[0, n, m, 0, p, 1].item()
                x0, y0 = int(np.floor(px)), int(np.floor(py))
                fx, fy = px - x0, py - y0
                acc = torch.zeros(32)
                for (yy, xx, wt) in ((y0, x0, (1 - fy) * (1 - fx)), (y0, x0 + 1, (1 - fy) * fx),
                                     (y0 + 1, x0, fy * (1 - fx)), (y0 + 1, x0 + 1, fy * fx)):
                    if 0 <= yy < h and 0 <= xx < w:
                        acc += wt * v[yy, xx, m]
                want[n, m * 32:(m + 1) * 32] += aw[0, n, m, 0, p] * acc
    assert torch.allclose(got[0], want, atol=2e-5)


def test_fp64_mode_close_to_fp32():
    cfg = O.OracleConfig(task="seg", num_classes=19, timesteps=2)
    W = O.make_weights(cfg, seed=4)
    x, noise = O.make_inputs(cfg, B=1, h=8, w=8, seed=6)
    o32 = O.sample(W, cfg, x, noise)
    o64 = O.sample(O.cast_weights(W, torch.float64), cfg, x.double(), noise.double())
    assert (o32.double() - o64).abs().max() < 1e-3


@pytest.mark.parametrize("path", golden_files("neck"), ids=lambda p: os.path.basename(p)[:-4])
def test_neck_oracle_matches_reference_output(path):
    """FPN + MultiStageMerging restatement vs the unmodified reference modules (make_golden.py neck)."""
    W, xs, g = load_neck_case(path)
    trace = {}
    out = NO.neck(W, xs, trace)
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    # same ATen ops in the same order on the same CPU: bit-for-bit
    assert torch.equal(out, ref), f"max |d| = {(out - ref).abs().max().item():.3e}"
    for i, o in enumerate(trace["fpn"]):
        assert torch.equal(o, torch.from_numpy(g[f"fpn{i}"])), f"fpn level {i}"


def test_neck_merge_commutes_with_the_1x1_conv():
    """The CUDA neck applies the 1x1 `down` conv per level BEFORE the bilinear resize (both are linear, the conv acts
    on channels, the resize on space): same result up to fp32 summation order."""
    import torch.nn.functional as F
    W = NO.make_weights([96, 192, 384, 768], seed=5)
    xs = NO.make_inputs([96, 192, 384, 768], 2, 9, 14, seed=6)
    outs = NO.fpn(W, xs)
    wd = W["neck.1.down.conv.weight"]
    pre = sum(F.interpolate(F.conv2d(o, wd[:, 256 * l:256 * (l + 1)]), size=outs[0].shape[2:], mode="bilinear",
                            align_corners=False) for l, o in enumerate(outs))
    got = F.group_norm(pre, NO.GROUPS, W["neck.1.down.gn.weight"], W["neck.1.down.gn.bias"], NO.EPS)
    want = NO.multi_stage_merging(W, outs)
    assert (got - want).abs().max().item() < 2e-5


@pytest.mark.parametrize("path", golden_files("bev"), ids=lambda p: os.path.basename(p)[:-4])
def test_bev_oracle_matches_reference_output(path):
    """BEV ddim_sample + head restatement vs the unmodified reference classes (make_golden.py bev)."""
    cfg, W, x, noise, g = load_bev_case(path)
    ref = torch.from_numpy(g["out"])
    trace = {}
    out = BO.ddim_sample_bev(W, cfg, x, noise, trace)
    assert out.shape == ref.shape
    d = (out - ref).abs().max().item()
    assert d < 2e-5, f"max |d| = {d:.3e}"          # plain tensors: matmul folding differs from nn.Parameter weights
    steps = torch.stack(trace["prob"])
    want = torch.from_numpy(g["step_prob"])
    assert (steps - want).abs().max().item() < 2e-5
    # the thresholded multi-hot maps (what feeds back into the state) agree everywhere outside fp32 ties at 0.5
    flips = (steps > cfg.threshold) != (want > cfg.threshold)
    assert int(flips.sum()) == 0 or float((want[flips] - cfg.threshold).abs().max()) < 1e-5
    Wp = {k: v.clone().requires_grad_(True) for k, v in W.items()}
    with torch.no_grad():
        out_p = BO.ddim_sample_bev(Wp, cfg, x, noise)
    assert torch.equal(out_p, ref), f"max |d| = {(out_p - ref).abs().max().item():.3e}"   # same ops, same order: bit-for-bit
