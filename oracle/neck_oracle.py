"""CPU oracle for the neck in front of the decode loop.  TEST INFRASTRUCTURE ONLY.

Restates, in plain PyTorch-CPU ops and without any mmcv / mmseg import, the two neck modules every shipped
DDP config chains in front of the decode head (``neck=[dict(type='FPN', ...), dict(type='MultiStageMerging', ...)]``,
e.g. segmentation/configs/cityscapes/ddp_swin_t_4x4_512x1024_160k_cityscapes.py:38-55 and
depth/configs/ddp_nyu/ddp_swint_1k_w7_nyu_bs2x8_scale01.py:40-56).  Nothing in the product path may import this file.

Parity status: PINNED.  ``tests/golden/make_golden.py neck`` builds the UNMODIFIED reference modules from the
reference's own config files (through ``refshim.py``), loads ``make_weights`` into them by state-dict key and stores
their outputs under ``tests/golden/neck_*.npz``; ``tests/test_oracle_golden.py`` checks this file against them.

Reference file:line followed (paths relative to /root/reference; "vmmcv" = controlnet/annotator/uniformer/mmcv):

  conv_module      vmmcv/cnn/bricks/conv_module.py:101-106 (bias='auto' -> no conv bias when a norm follows),
                   :193-205 (order conv -> norm -> act; act_cfg=None in every DDP config)
  fpn              segmentation/mmseg/models/necks/fpn.py:119-140 (lateral 1x1 / fpn 3x3 ConvModules),
                   :162-213 (forward: laterals, top-down `resize(..., mode='nearest')` adds, 3x3 outputs)
  multi_stage_merging  segmentation/mmseg/models/necks/multi_stage_merging.py:40-52 (bilinear resize of every level to
                   level 0's size, align_corners=False, concat, 1x1 ConvModule with GN)
  resize           segmentation/mmseg/ops/wrappers.py:8-27 (F.interpolate)
  group norm       vmmcv/cnn/bricks/norm.py:72-107 (nn.GroupNorm(num_groups, C), eps 1e-5)

The depth tree's copies (depth/depth/models/necks/fpn.py, multi_stage_merging.py) differ from the segmentation
files only in their import lines.
"""
import math
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

OUT = 256
GROUPS = 32
EPS = 1e-5


def make_weights(in_channels: Sequence[int], seed: int = 0, out_channels: int = OUT) -> Dict[str, torch.Tensor]:
    """Seeded weights under the reference's state-dict keys (``neck.0.*`` = FPN, ``neck.1.*`` = MultiStageMerging).

    Convolutions are xavier-uniform as the reference's init_cfg asks (fpn.py:82-83); the GroupNorm affine
    parameters are perturbed from (1, 0) so that a swapped gamma / beta or a wrong group shows up."""
    g = torch.Generator().manual_seed(seed)

    def xavier(*shape):
        rf = shape[2] * shape[3]
        a = math.sqrt(6.0 / (shape[1] * rf + shape[0] * rf))
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    W: Dict[str, torch.Tensor] = {}

    def gn(prefix):
        W[prefix + "gn.weight"] = 1.0 + 0.1 * torch.randn(out_channels, generator=g)
        W[prefix + "gn.bias"] = 0.1 * torch.randn(out_channels, generator=g)

    for i, c in enumerate(in_channels):
        W[f"neck.0.lateral_convs.{i}.conv.weight"] = xavier(out_channels, c, 1, 1)
        gn(f"neck.0.lateral_convs.{i}.")
    for i in range(len(in_channels)):
        W[f"neck.0.fpn_convs.{i}.conv.weight"] = xavier(out_channels, out_channels, 3, 3)
        gn(f"neck.0.fpn_convs.{i}.")
    W["neck.1.down.conv.weight"] = xavier(out_channels, out_channels * len(in_channels), 1, 1)
    gn("neck.1.down.")
    return W


def make_inputs(in_channels: Sequence[int], B: int, h0: int, w0: int, seed: int = 0) -> List[torch.Tensor]:
    """Backbone pyramid stand-in: level l is (B, C_l, ceil(h0 / 2^l), ceil(w0 / 2^l)), N(0, 1)."""
    g = torch.Generator().manual_seed(seed)
    xs, h, w = [], h0, w0
    for c in in_channels:
        xs.append(torch.randn(B, c, h, w, generator=g))
        h, w = (h + 1) // 2, (w + 1) // 2
    return xs


def conv_module(W, prefix, x, padding=0):
    """ConvModule(conv -> GN), no conv bias, no activation (conv_module.py:101-106, 193-205)."""
    y = F.conv2d(x, W[prefix + "conv.weight"], None, padding=padding)
    return F.group_norm(y, GROUPS, W[prefix + "gn.weight"], W[prefix + "gn.bias"], EPS)


def fpn(W, inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    """FPN.forward (fpn.py:162-213) for num_outs == len(inputs), start_level 0, upsample nearest."""
    L = len(inputs)
    lat = [conv_module(W, f"neck.0.lateral_convs.{i}.", inputs[i]) for i in range(L)]
    for i in range(L - 1, 0, -1):                                             # fpn.py:173-183
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    return [conv_module(W, f"neck.0.fpn_convs.{i}.", lat[i], padding=1) for i in range(L)]


def multi_stage_merging(W, inputs: Sequence[torch.Tensor]) -> torch.Tensor:
    """MultiStageMerging.forward (multi_stage_merging.py:40-52)."""
    size = inputs[0].shape[2:]
    ups = [F.interpolate(x, size=size, mode="bilinear", align_corners=False) for x in inputs]
    return conv_module(W, "neck.1.down.", torch.cat(ups, dim=1))


def neck(W, inputs: Sequence[torch.Tensor], trace: dict = None) -> torch.Tensor:
    """The whole neck: backbone pyramid -> x (B, 256, h0, w0), the frozen conditioning feature of the decode loop."""
    outs = fpn(W, inputs)
    if trace is not None:
        trace["fpn"] = outs
    return multi_stage_merging(W, outs)
