"""Helpers shared by the golden-fixture tests (regenerate inputs from stored seeds)."""
import glob
import os

import numpy as np
import torch

from oracle import ddp_oracle as O
from oracle import neck_oracle as NO
from oracle import bev_oracle as BO

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


LOOP_TASKS = ("seg", "depth")          # fixtures of the decode loop; "neck" fixtures are loaded by load_neck_case


def golden_files(task=None):
    """Fixtures whose file name starts with `task`; default: every decode-loop fixture (seg_*, depth_*)."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(f)
        if name.startswith(task if task is not None else LOOP_TASKS):
            out.append(f)
    return out


def checksum(t):
    t = t.double()
    return np.array([float(t.sum()), float(t.abs().sum())])


def load_case(path):
    """-> (cfg, W, x (1,256,h,w), noise (1,R,Cin,h,w), golden dict)."""
    g = dict(np.load(path, allow_pickle=False))
    task = str(g["task"])
    if task == "seg":
        cfg = O.OracleConfig(task="seg", num_classes=int(g["num_classes"]), timesteps=int(g["timesteps"]),
                             randsteps=int(g["randsteps"]), bit_scale=float(g["bit_scale"]),
                             accumulation=bool(g["accumulation"]),
                             diffusion=str(g["diffusion"]) if "diffusion" in g else "ddim")
        cin = 256
    else:
        cfg = O.OracleConfig(task="depth", timesteps=int(g["timesteps"]), randsteps=int(g["randsteps"]),
                             bit_scale=float(g["bit_scale"]), min_depth=float(g["min_depth"]),
                             max_depth=float(g["max_depth"]))
        cin = 1
    W = O.make_weights(cfg, seed=int(g["wseed"]))
    h, w, R = int(g["h"]), int(g["w"]), cfg.randsteps
    x = torch.randn(1, 256, h, w, generator=torch.Generator().manual_seed(int(g["xseed"])))
    state = torch.get_rng_state()
    torch.manual_seed(int(g["nseed"]))
    noise = torch.randn((R, cin, h, w))[None]
    if cfg.diffusion == "ddpm":       # the reference then draws randn_like(mask_t) once per step (ddp.py:279-283)
        g["ddpm_noise"] = torch.stack([torch.randn((R, cin, h, w)) for _ in range(cfg.timesteps)])[None]   # (1,T,R,C,h,w)
    torch.set_rng_state(state)
    # the fixtures were produced from exactly these tensors
    assert np.allclose(checksum(x), g["x_checksum"], rtol=0, atol=1e-6), "x regeneration drifted"
    assert np.allclose(checksum(noise), g["noise_checksum"], rtol=0, atol=1e-6), "noise regeneration drifted"
    wsum = checksum(torch.cat([v.flatten() for _, v in sorted(W.items())]))
    assert np.allclose(wsum, g["w_checksum"], rtol=0, atol=1e-5), "weight regeneration drifted"
    return cfg, W, x, noise, g


def load_neck_case(path):
    """-> (W, inputs [4 x (B,C_l,h_l,w_l)], golden dict) of a neck_*.npz fixture."""
    g = dict(np.load(path, allow_pickle=False))
    assert str(g["task"]) == "neck"
    in_channels = [int(c) for c in g["in_channels"]]
    W = NO.make_weights(in_channels, seed=int(g["wseed"]))
    xs = NO.make_inputs(in_channels, int(g["B"]), int(g["h"]), int(g["w"]), seed=int(g["xseed"]))
    assert np.allclose(checksum(torch.cat([x.flatten() for x in xs])), g["x_checksum"], rtol=0, atol=1e-6)
    wsum = checksum(torch.cat([v.flatten() for _, v in sorted(W.items())]))
    assert np.allclose(wsum, g["w_checksum"], rtol=0, atol=1e-5), "weight regeneration drifted"
    return W, xs, g


def load_bev_case(path):
    """-> (cfg, W, x (1,feat,h,w), noise (R,256,h,w), golden dict) of a bev_*.npz fixture."""
    g = dict(np.load(path, allow_pickle=False))
    assert str(g["task"]) == "bev"
    cfg = BO.BevConfig(timesteps=int(g["timesteps"]), randsteps=int(g["randsteps"]), bit_scale=float(g["bit_scale"]),
                       num_layers=int(g["num_layers"]), feat_channels=int(g["feat_channels"]),
                       input_scope=tuple(tuple(float(v) for v in r) for r in g["input_scope"]),
                       output_scope=tuple(tuple(float(v) for v in r) for r in g["output_scope"]))
    W = BO.make_weights(cfg, seed=int(g["wseed"]))
    h, w = int(g["h"]), int(g["w"])
    x = torch.randn(1, cfg.feat_channels, h, w, generator=torch.Generator().manual_seed(int(g["xseed"])))
    state = torch.get_rng_state()
    torch.manual_seed(int(g["nseed"]))
    noise = torch.randn((cfg.randsteps, 256, h, w))
    torch.set_rng_state(state)
    assert np.allclose(checksum(x), g["x_checksum"], rtol=0, atol=1e-6), "x regeneration drifted"
    assert np.allclose(checksum(noise), g["noise_checksum"], rtol=0, atol=1e-6), "noise regeneration drifted"
    wsum = checksum(torch.cat([v.flatten() for _, v in sorted(W.items())]))
    assert np.allclose(wsum, g["w_checksum"], rtol=0, atol=1e-5), "weight regeneration drifted"
    return cfg, W, x, noise, g
