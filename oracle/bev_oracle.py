"""CPU oracle for the BEV map-segmentation variant of the decode loop.  TEST INFRASTRUCTURE ONLY.

Restates, in plain PyTorch-CPU ops, ``DDP.ddim_sample`` of the BEV tree and the head it calls.  Same denoiser layers
as the segmentation path (the shared functions of ``ddp_oracle`` are reused), different boundary work around them:
the head resamples its input onto the output BEV grid (``BEVGridTransform``: 128x128 -> 200x200 in the shipped configs),
has 5 layers, ends in a 6-channel ``sigmoid``; the loop thresholds the sigmoid maps at 0.5, nearest-resizes the multi-hot
map back to the state grid, embeds every class slot and takes the MEAN of the six embeddings, and returns the mean of
the sigmoid maps of ALL steps and samples.  Nothing in the product path may import this file.

Parity status: PINNED.  ``tests/golden/make_golden.py bev`` runs the UNMODIFIED reference classes
(bev/mmdet3d/models/fusion_models/ddp.py::DDP.ddim_sample with bev/mmdet3d/models/heads/segm/
deformable_head_with_time.py::DeformableHeadWithTime, head arguments from bev/configs/nuscenes/seg/*.yaml) and
``tests/test_oracle_golden.py`` compares this file with the stored outputs.

Reference file:line followed (paths relative to /root/reference/bev):

  time pairs          mmdet3d/models/fusion_models/ddp.py:128-136 (= the segmentation pairs with sample_range[0] = 0)
  log_snr, alpha/sigma, time_mlp   mmdet3d/models/fusion_models/ddp.py:31-60, 97-116 (same formulas as segmentation)
  grid_transform      mmdet3d/models/heads/segm/deformable_head_with_time.py:58-98
  head_bev            mmdet3d/models/heads/segm/deformable_head_with_time.py:178-241 (sigmoid at :241)
  ddim_sample_bev     mmdet3d/models/fusion_models/ddp.py:268-301

State-dict keys: ``embedding_table.weight`` (7, 256), ``transform.conv.{weight,bias}`` (256, feat_channels + 256, 1, 1), ``time_mlp.*``,
``heads.map.encoder.layers.N.*``, ``heads.map.conv_seg.{weight,bias}``.
"""
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import ddp_oracle as O

NUM_CLASSES = 6          # hard-coded in the reference (fusion_models/ddp.py:89)


@dataclass
class BevConfig:
    timesteps: int = 3
    randsteps: int = 5
    time_difference: int = 1
    bit_scale: float = 0.01
    threshold: float = 0.5
    num_layers: int = 5
    feat_channels: int = 256      # channels of the fused BEV feature x: 256 (camera-only config) or 512 (default, fusion config)
    input_scope: Tuple = ((-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8))
    output_scope: Tuple = ((-50.0, 50.0, 0.5), (-50.0, 50.0, 0.5))

    def seg_config(self) -> O.OracleConfig:
        return O.OracleConfig(task="seg", num_classes=NUM_CLASSES, timesteps=self.timesteps, randsteps=self.randsteps,
                              time_difference=self.time_difference, bit_scale=self.bit_scale,
                              num_layers=self.num_layers)


def make_weights(cfg: BevConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """The segmentation recipe under the BEV tree's keys (decode_head.* -> heads.map.*)."""
    W = O.make_weights(cfg.seg_config(), seed=seed)
    if cfg.feat_channels != O.E:      # transform = ConvModule(tmp_channels + feat_channels, tmp_channels, 1) (ddp.py:104)
        g = torch.Generator().manual_seed(seed + 7919)
        a = (6.0 / (cfg.feat_channels + 2 * O.E)) ** 0.5
        W["transform.conv.weight"] = (torch.rand(O.E, cfg.feat_channels + O.E, 1, 1, generator=g) * 2 - 1) * a
    return {("heads.map." + k[len("decode_head."):] if k.startswith("decode_head.") else k): v for k, v in W.items()}


def _as_seg_keys(W):
    return {("decode_head." + k[len("heads.map."):] if k.startswith("heads.map.") else k): v for k, v in W.items()}


def grid_transform(x, input_scope, output_scope):
    """BEVGridTransform.forward with prescale_factor 1 (deformable_head_with_time.py:72-98)."""
    coords = []
    for (imin, imax, _), (omin, omax, ostep) in zip(input_scope, output_scope):
        v = torch.arange(omin + ostep / 2, omax, ostep)
        v = (v - imin) / (imax - imin) * 2 - 1
        coords.append(v)
    u, v = torch.meshgrid(coords, indexing="ij")
    grid = torch.stack([v, u], dim=-1)
    grid = torch.stack([grid] * x.shape[0], dim=0)
    return F.grid_sample(x, grid, mode="bilinear", align_corners=False)


def head_bev(W, cfg: BevConfig, feat, time, trace=None):
    """feat (rows, 256, h, w), time (1, 1024) -> sigmoid maps (rows, 6, H', W') on the output grid."""
    Ws = _as_seg_keys(W)
    feat = grid_transform(feat, cfg.input_scope, cfg.output_scope)
    mem = O.head_tokens(Ws, cfg.seg_config(), feat, time)
    x = F.conv2d(mem, Ws["decode_head.conv_seg.weight"], Ws["decode_head.conv_seg.bias"])
    if trace is not None:
        trace.setdefault("feat_grid", []).append(feat)
        trace.setdefault("logit", []).append(x)
    return torch.sigmoid(x)


def grid_coords(input_scope, output_scope):
    """The normalised sampling coordinates of BEVGridTransform (rows, columns) — deformable_head_with_time.py:82-88."""
    coords = []
    for (imin, imax, _), (omin, omax, ostep) in zip(input_scope, output_scope):
        v = torch.arange(omin + ostep / 2, omax, ostep)
        coords.append((v - imin) / (imax - imin) * 2 - 1)
    return coords


def ddim_sample_bev(W, cfg: BevConfig, x, noise, trace=None):
    """x (1, feat_channels, h, w); noise (R, 256, h, w) -> (1, 6, H', W'): mean over all T * R sigmoid maps (ddp.py:268-301)."""
    assert x.shape[0] == 1, "the reference loop is defined for one sample (its batch axis is randsteps)"
    R = cfg.randsteps
    h, w = x.shape[2:]
    xr = x.repeat(R, 1, 1, 1)
    mask_t = noise
    outs = []
    for t_now, t_next in O.time_pairs_seg(cfg.seg_config()):
        tn = torch.tensor([t_now], dtype=torch.float32)
        tx = torch.tensor([t_next], dtype=torch.float32)
        feat = F.conv2d(torch.cat([xr, mask_t], dim=1), W["transform.conv.weight"], W["transform.conv.bias"])
        log_snr, log_snr_next = O.log_snr_cosine(tn), O.log_snr_cosine(tx)
        alpha, sigma = O.alpha_sigma(log_snr.view(1, 1, 1, 1))
        alpha_next, sigma_next = O.alpha_sigma(log_snr_next.view(1, 1, 1, 1))
        temb = O.time_mlp(W, log_snr)
        prob = head_bev(W, cfg, feat, temb, trace)
        pred = (prob > cfg.threshold)
        factor = (torch.arange(NUM_CLASSES) + 1).view(1, NUM_CLASSES, 1, 1)
        pred = pred * factor
        pred = F.interpolate(pred.float(), size=(h, w), mode="nearest").to(torch.int64)
        pred = F.embedding(pred, W["embedding_table.weight"]).mean(dim=1).permute(0, 3, 1, 2)
        pred = (torch.sigmoid(pred) * 2 - 1) * cfg.bit_scale
        eps = (mask_t - alpha * pred) / sigma.clamp(min=1e-8)
        mask_t = pred * alpha_next + eps * sigma_next
        outs.append(prob)
        if trace is not None:
            trace.setdefault("feat", []).append(feat)
            trace.setdefault("prob", []).append(prob)
            trace.setdefault("mask_t", []).append(mask_t)
    return torch.cat(outs, dim=0).mean(dim=0, keepdim=True)


def sample(W, cfg: BevConfig, x, noise):
    """Batched generalisation: x (B, 256, h, w), noise (B, R, 256, h, w) -> (B, 6, H', W'); per-image reference loops."""
    return torch.cat([ddim_sample_bev(W, cfg, x[b:b + 1], noise[b]) for b in range(x.shape[0])], dim=0)
