"""Host emulation of the neck's CUDA launch sequence vs the oracle and the reference goldens (no GPU needed).

ddp_b200/csrc/neck_plan.h holds the neck's launch sequence, per-element kernel bodies, weight repack and workspace
carve-up as backend-agnostic C++; tests/emu/neck_emu.cpp runs exactly that code with a sequential host backend.
This is a CHECK of the product's indexing and arithmetic in a GPU-less container, not a product path: the library
itself has no CPU implementation.  The CUDA build of the same sequence is covered by tests/test_gpu_neck.py.
"""
import ctypes
import os
import subprocess

import pytest
import torch

from oracle import neck_oracle as NO
from golden_util import golden_files, load_neck_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("neck_emu") / "libneck_emu.so")
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "ddp_b200", "csrc"),
           os.path.join(ROOT, "tests", "emu", "neck_emu.cpp"), "-o", out]
    subprocess.run(cmd, check=True)
    lib = ctypes.CDLL(out)
    lib.neck_emu_forward.restype = ctypes.c_int
    return lib


def _ptrs(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def run_emu(lib, W, xs, stages=3, want_fpn=True):
    L, B = len(xs), xs[0].shape[0]
    xs = [x.contiguous() for x in xs]
    C = (ctypes.c_int * L)(*[x.shape[1] for x in xs])
    H = (ctypes.c_int * L)(*[x.shape[2] for x in xs])
    Wd = (ctypes.c_int * L)(*[x.shape[3] for x in xs])
    lat_w = [W[f"neck.0.lateral_convs.{l}.conv.weight"].contiguous() for l in range(L)]
    fpn_w = [W[f"neck.0.fpn_convs.{l}.conv.weight"].contiguous() for l in range(L)]
    down_w = W["neck.1.down.conv.weight"].contiguous()
    gn = []
    for stem in ("neck.0.lateral_convs", "neck.0.fpn_convs"):
        for l in range(L):
            gn += [W[f"{stem}.{l}.gn.weight"].contiguous(), W[f"{stem}.{l}.gn.bias"].contiguous()]
    gn += [W["neck.1.down.gn.weight"].contiguous(), W["neck.1.down.gn.bias"].contiguous()]
    x_out = torch.empty(B, 256, xs[0].shape[2], xs[0].shape[3])
    fpn_outs = [torch.empty(B, 256, x.shape[2], x.shape[3]) for x in xs] if want_fpn else [None] * L
    rc = lib.neck_emu_forward(stages, L, B, C, H, Wd, 32, ctypes.c_float(1e-5), _ptrs(lat_w), _ptrs(fpn_w),
                              ctypes.c_void_p(down_w.data_ptr()), _ptrs(gn), _ptrs(xs),
                              ctypes.c_void_p(x_out.data_ptr()), _ptrs(fpn_outs))
    assert rc == 0
    return x_out, fpn_outs


TOL = 5e-5     # fp32 summation order (K up to 2304) on O(1) GroupNorm outputs


@pytest.mark.parametrize("path", golden_files("neck"), ids=lambda p: os.path.basename(p)[:-4])
def test_emulated_neck_matches_reference_golden(emu, path):
    W, xs, g = load_neck_case(path)
    x, fpn = run_emu(emu, W, xs)
    for l, o in enumerate(fpn):
        d = (o - torch.from_numpy(g[f"fpn{l}"])).abs().max().item()
        assert d < TOL, f"fpn level {l}: max|d| = {d:.3e}"
    d = (x - torch.from_numpy(g["out"])).abs().max().item()
    assert d < TOL, f"x: max|d| = {d:.3e}"


@pytest.mark.parametrize("B,h,w", [(1, 1, 1), (2, 5, 3), (1, 9, 17), (3, 4, 4)])
def test_emulated_neck_matches_oracle_on_ragged_shapes(emu, B, h, w):
    chans = [96, 192, 384, 768]
    W = NO.make_weights(chans, seed=h * 31 + w)
    xs = NO.make_inputs(chans, B, h, w, seed=7)
    trace = {}
    want = NO.neck(W, xs, trace)
    x, fpn = run_emu(emu, W, xs)
    for l, o in enumerate(fpn):
        assert (o - trace["fpn"][l]).abs().max().item() < TOL
    assert (x - want).abs().max().item() < TOL


def test_emulated_stage_split_equals_fused(emu):
    """FPN-only followed by merge-only (the standalone FPN / MultiStageMerging modules) == the fused sequence."""
    chans = [96, 192, 384, 768]
    W = NO.make_weights(chans, seed=9)
    xs = NO.make_inputs(chans, 2, 6, 10, seed=8)
    fused, fpn_fused = run_emu(emu, W, xs, stages=3)
    _, fpn_only = run_emu(emu, W, xs, stages=1)
    for a, b in zip(fpn_fused, fpn_only):
        assert torch.equal(a, b)
    merged, _ = run_emu(emu, W, fpn_only, stages=2, want_fpn=False)
    assert (merged - fused).abs().max().item() < 1e-6
