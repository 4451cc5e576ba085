"""Host emulation of the BEV loop's CUDA launch sequence vs the oracle and the reference goldens (no GPU needed).

ddp_b200/csrc/bev_plan.h holds what the BEV variant adds around the (shared, hardware-verified) denoiser: the
feat_channels-wide transform, the grid_sample onto the output grid, sigmoid accumulation, threshold -> nearest resize ->
mean class embedding -> DDIM update.  tests/emu/bev_emu.cpp runs exactly that code on the CPU with the denoiser
replaced by a replay of the oracle's per-step logits (teacher forcing).  A CHECK of the product's indexing and
arithmetic, not a product path.  The CUDA build is covered by tests/test_zzz_gpu_bev.py.
"""
import ctypes
import os
import subprocess

import pytest
import torch

from ddp_b200 import schedule as S
from oracle import bev_oracle as BO
from golden_util import golden_files, load_bev_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("bev_emu") / "libbev_emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "ddp_b200", "csrc"),
                    os.path.join(ROOT, "tests", "emu", "bev_emu.cpp"), "-o", out], check=True)
    lib = ctypes.CDLL(out)
    lib.bev_emu_run.restype = ctypes.c_int
    return lib


def run_emu(lib, cfg, W, x, noise, logits):
    """x (B,feat,h,w), noise (B,R,256,h,w), logits [T] x (B*R,6,Ho,Wo) -> (out, feat dumps [T], state dumps [T])."""
    B, feat, h, w = x.shape
    R, T = cfg.randsteps, cfg.timesteps
    gy, gx = [c.contiguous() for c in BO.grid_coords(cfg.input_scope, cfg.output_scope)]
    Ho, Wo = gy.numel(), gx.numel()
    l, a, s, an, sn = S.seg_schedule(T, cfg.time_difference, (0, 0.999), "cosine")
    sched = torch.tensor([a, s, an, sn], dtype=torch.float32).contiguous()
    replay = torch.stack(logits).contiguous()
    out = torch.empty(B, 6, Ho, Wo)
    feat_dump = torch.empty(T, B * R, 256, Ho, Wo)
    state_dump = torch.empty(T, B * R, h * w, 256)
    tw, tb = W["transform.conv.weight"].contiguous(), W["transform.conv.bias"].contiguous()
    emb = W["embedding_table.weight"].contiguous()
    x, noise = x.contiguous(), noise.contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    f = ctypes.c_float
    rc = lib.bev_emu_run(B, R, T, feat, h, w, Ho, Wo, f(cfg.bit_scale), f(cfg.threshold), p(tw), p(tb), p(emb), p(gy), p(gx),
                         p(sched), p(x), p(noise), p(replay), p(out), p(feat_dump), p(state_dump))
    assert rc == 0
    return out, feat_dump, state_dump


def check(lib, cfg, W, x, noise_b, ref_out=None):
    """Teacher-forced comparison for every image of the batch against per-image oracle runs."""
    B = x.shape[0]
    traces, outs = [], []
    for b in range(B):
        tr = {}
        outs.append(BO.ddim_sample_bev(W, cfg, x[b:b + 1], noise_b[b], tr))
        traces.append(tr)
    logits = [torch.cat([traces[b]["logit"][k] for b in range(B)]) for k in range(cfg.timesteps)]
    out, feat_dump, state_dump = run_emu(lib, cfg, W, x, noise_b, logits)
    want = torch.cat(outs)
    assert (out - want).abs().max().item() < 2e-6, "mean of the sigmoid maps"
    if ref_out is not None:
        assert (out - ref_out).abs().max().item() < 2e-5
    R = cfg.randsteps
    for k in range(cfg.timesteps):
        fg = torch.cat([traces[b]["feat_grid"][k] for b in range(B)])
        d = (feat_dump[k] - fg).abs().max().item()
        assert d < 2e-5, f"step {k}: denoiser input (transform + grid_sample) off by {d:.3e}"
        st = torch.cat([traces[b]["mask_t"][k] for b in range(B)])                     # (B*R, 256, h, w)
        got = state_dump[k].view(B * R, x.shape[2], x.shape[3], 256).permute(0, 3, 1, 2)
        d = (got - st).abs().max().item()
        assert d < 1e-5, f"step {k}: state after the DDIM update off by {d:.3e}"


@pytest.mark.parametrize("path", golden_files("bev"), ids=lambda p: os.path.basename(p)[:-4])
def test_emulated_bev_loop_matches_reference_golden(emu, path):
    cfg, W, x, noise, g = load_bev_case(path)
    check(emu, cfg, W, x, noise[None], torch.from_numpy(g["out"]))


def test_emulated_bev_loop_batched_and_out_of_range_grid(emu):
    """Two images, and an output scope that reaches outside the input scope (zero padding of grid_sample)."""
    cfg = BO.BevConfig(timesteps=2, randsteps=2, feat_channels=512, num_layers=1,
                       input_scope=((-4.0, 4.0, 0.8), (-4.8, 4.8, 0.8)), output_scope=((-5.0, 5.0, 0.5), (-4.0, 6.0, 0.5)))
    W = BO.make_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 512, 10, 12, generator=g)
    noise = torch.randn(2, 2, 256, 10, 12, generator=g)
    check(emu, cfg, W, x, noise)
