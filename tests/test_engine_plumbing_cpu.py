"""Dry run of the Python plumbing of NeckEngine / BevDecodeEngine without a GPU: the ctypes library is replaced by a
recorder and torch.cuda's stream / device context by dummies, so every line between the public call and the C ABI call
executes on CPU tensors (shape checks, pointer arrays, workspace sizing and alignment, argument order).  What the
library DOES with the arguments is covered by the emulation tests (kernel bodies) and the GPU tests."""
import contextlib
import ctypes

import pytest
import torch

from ddp_b200 import _lib as L
from ddp_b200.bev import BevDecodeEngine, grid_coords
from ddp_b200.neck import NeckEngine


class Recorder:
    """Stands in for the CDLL: every ddp_* function records its arguments and returns 0."""

    def __init__(self, ws_bytes=4096):
        self.calls, self.ws_bytes = [], ws_bytes

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            if name.endswith("_plan"):
                args[-1]._obj.value = self.ws_bytes          # byref(c_size_t)
            return 0
        return fn


@pytest.fixture
def fake_cuda(monkeypatch):
    class _Stream:
        cuda_stream = 0xABC
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: _Stream())
    monkeypatch.setattr(torch.cuda, "device", lambda device=None: contextlib.nullcontext())


def test_neck_engine_forward_plumbing(fake_cuda):
    eng = NeckEngine.__new__(NeckEngine)
    eng.lib, eng.device, eng._h = Recorder(), torch.device("cpu"), ctypes.c_void_p(1)
    eng.in_channels, eng.stages, eng.levels = [96, 192, 384, 768], 3, 4
    eng._plan = eng._ws = None
    xs = [torch.randn(2, c, 8 >> l, 12 >> l) for l, c in enumerate(eng.in_channels)]
    x, fpn = eng.forward(xs, want_fpn=True)
    assert tuple(x.shape) == (2, 256, 8, 12) and [tuple(t.shape) for t in fpn] == [(2, 256, 8 >> l, 12 >> l) for l in range(4)]
    (n1, a1), (n2, a2) = eng.lib.calls
    assert n1 == "ddp_neck_plan" and a1[1] == 2 and list(a1[2]) == [8, 4, 2, 1] and list(a1[3]) == [12, 6, 3, 1]
    assert n2 == "ddp_neck_forward"
    assert [p for p in a2[1]] == [t.data_ptr() for t in xs]              # contiguous fp32 inputs are passed as they are
    assert a2[2].value == x.data_ptr() and [p for p in a2[3]] == [t.data_ptr() for t in fpn]
    assert a2[4].value % 256 == 0 and a2[5] == 4096 and a2[6].value == 0xABC
    assert eng._ws.data_ptr() <= a2[4].value and a2[4].value + a2[5] <= eng._ws.data_ptr() + eng._ws.numel()
    eng.forward(xs)                                                       # same geometry: no second plan
    assert [n for n, _ in eng.lib.calls].count("ddp_neck_plan") == 1
    assert eng.lib.calls[-1][1][3] is None                                # fused neck without FPN copies: NULL array
    with pytest.raises(ValueError, match="expected"):
        eng.forward([xs[0], xs[1], xs[2], torch.randn(2, 512, 1, 1)])
    with pytest.raises(AssertionError):
        eng.forward(xs[:3])
    # FPN-only stage: no x_out, the FPN outputs are mandatory
    eng.stages, eng._plan = L.NECK_STAGE_FPN, None
    x, fpn = eng.forward(xs)
    assert x is None and len(fpn) == 4 and eng.lib.calls[-1][1][2] is None
    # merge-only stage: 256-channel inputs
    eng.stages, eng._plan = L.NECK_STAGE_MERGE, None
    x, fpn = eng.forward([torch.randn(1, 256, 4 >> l, 4 >> l) for l in range(3)] + [torch.randn(1, 256, 1, 1)])
    assert tuple(x.shape) == (1, 256, 4, 4) and fpn is None


def test_bev_engine_sample_plumbing(fake_cuda):
    eng = BevDecodeEngine.__new__(BevDecodeEngine)
    eng.lib, eng.device, eng._h = Recorder(ws_bytes=1 << 16), torch.device("cpu"), ctypes.c_void_p(1)
    eng.feat_channels, eng.timesteps = 512, 3
    eng._plan = eng._ws = None
    gy, gx = grid_coords(((-4.8, 4.8, 0.8), (-4.0, 4.0, 0.8)), ((-4.5, 4.5, 0.5), (-3.5, 3.5, 0.5)))
    x, noise = torch.randn(2, 512, 12, 10), torch.randn(2, 5, 256, 12, 10)
    out = eng.sample(x, noise, gy, gx)
    assert tuple(out.shape) == (2, 6, 18, 14)
    (n1, a1), (n2, a2) = eng.lib.calls
    assert n1 == "ddp_bev_plan" and a1[1:7] == (2, 5, 12, 10, 18, 14)
    assert [a1[7][i] for i in range(18)] == gy.tolist() and [a1[8][i] for i in range(14)] == gx.tolist()
    assert n2 == "ddp_bev_sample" and a2[1].value == x.data_ptr() and a2[2].value == noise.data_ptr()
    assert a2[3].value == out.data_ptr() and a2[4].value % 256 == 0 and a2[5] == 1 << 16 and a2[6].value == 0xABC
    eng.sample(x, noise, gy, gx)
    assert [n for n, _ in eng.lib.calls].count("ddp_bev_plan") == 1
    with pytest.raises(ValueError, match="expected"):
        eng.sample(torch.randn(2, 256, 12, 10), noise, gy, gx)
    with pytest.raises(ValueError, match="noise"):
        eng.sample(x, torch.randn(2, 5, 256, 12, 11), gy, gx)
    assert tuple(eng.sample(x[:0], noise[:0], gy, gx).shape) == (0, 6, 18, 14)
