"""Seeded synthetic weights and inputs of the reference's shapes for bench.py and
smoke(): reference state-dict keys, non-degenerate deformable-offset / attention weights.  (tests/ checks
that this recipe generates the same tensors as the oracle's.)"""
import math
from typing import Dict

import torch

E, HEADS, POINTS, FFN_DIM, TIME_DIM = 256, 8, 4, 1024, 1024


def make_weights(task="seg", num_classes=19, learned_sinusoidal_dim=16, num_layers=6, seed=0) -> Dict[str, torch.Tensor]:
    """Seeded, non-degenerate weights with the reference's state-dict keys.

    sampling_offsets keep the reference's ring bias
    (vmmcv/ops/multi_scale_deform_attn.py:233-244) but get a non-zero weight so
    that the gather is data-dependent; attention_weights are xavier.
    """
    g = torch.Generator().manual_seed(seed)

    def xavier(*shape):
        fan_out, fan_in = shape[0], int(torch.tensor(shape[1:]).prod())
        a = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(*shape, generator=g) * 2 - 1) * a

    def ubias(n, a=0.1):
        return (torch.rand(n, generator=g) * 2 - 1) * a

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    W: Dict[str, torch.Tensor] = {}
    cin = E if task == "seg" else 1
    if task == "seg":
        W["embedding_table.weight"] = randn(num_classes + 1, E)
        W["transform.conv.weight"] = xavier(E, 2 * E, 1, 1)
        W["transform.conv.bias"] = ubias(E)
    else:
        W["down.conv.weight"] = xavier(E, E + cin, 1, 1)
        W["down.conv.bias"] = ubias(E)
    W["time_mlp.0.weights"] = randn(learned_sinusoidal_dim // 2)
    W["time_mlp.1.weight"] = xavier(TIME_DIM, learned_sinusoidal_dim + 1)
    W["time_mlp.1.bias"] = ubias(TIME_DIM)
    W["time_mlp.3.weight"] = xavier(TIME_DIM, TIME_DIM)
    W["time_mlp.3.bias"] = ubias(TIME_DIM)
    thetas = torch.arange(HEADS, dtype=torch.float32) * (2.0 * math.pi / HEADS)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(HEADS, 1, 1, 2).repeat(1, 1, POINTS, 1)
    for i in range(POINTS):
        grid[:, :, i, :] *= i + 1
    for j in range(num_layers):
        p = f"decode_head.encoder.layers.{j}."
        W[p + "attentions.0.sampling_offsets.weight"] = randn(HEADS * POINTS * 2, E, std=0.5 / 16.0)
        W[p + "attentions.0.sampling_offsets.bias"] = grid.reshape(-1).clone()
        W[p + "attentions.0.attention_weights.weight"] = xavier(HEADS * POINTS, E)
        W[p + "attentions.0.attention_weights.bias"] = ubias(HEADS * POINTS)
        W[p + "attentions.0.value_proj.weight"] = xavier(E, E)
        W[p + "attentions.0.value_proj.bias"] = ubias(E)
        W[p + "attentions.0.output_proj.weight"] = xavier(E, E)
        W[p + "attentions.0.output_proj.bias"] = ubias(E)
        W[p + "time_mlp.1.weight"] = xavier(2 * E, TIME_DIM)
        W[p + "time_mlp.1.bias"] = ubias(2 * E)
        W[p + "ffns.0.layers.0.0.weight"] = xavier(FFN_DIM, E)
        W[p + "ffns.0.layers.0.0.bias"] = ubias(FFN_DIM)
        W[p + "ffns.0.layers.1.weight"] = xavier(E, FFN_DIM)
        W[p + "ffns.0.layers.1.bias"] = ubias(E)
        for k in (0, 1):
            W[p + f"norms.{k}.weight"] = 1.0 + 0.1 * randn(E)
            W[p + f"norms.{k}.bias"] = 0.1 * randn(E)
    if task == "seg":
        W["decode_head.conv_seg.weight"] = xavier(num_classes, E, 1, 1)
        W["decode_head.conv_seg.bias"] = ubias(num_classes)
    else:
        W["decode_head.conv_depth.weight"] = xavier(1, E, 3, 3)
        W["decode_head.conv_depth.bias"] = torch.full((1,), 3.0)
    return W



def make_inputs(task, randsteps, B, h, w, seed=1234):
    g = torch.Generator().manual_seed(seed)
    cin = E if task == "seg" else 1
    x = torch.randn(B, E, h, w, generator=g)
    noise = torch.randn(B, randsteps, cin, h, w, generator=g)
    return x, noise


def make_neck_weights(in_channels, seed=0, out_channels=E) -> Dict[str, torch.Tensor]:
    """FPN + MultiStageMerging weights under the reference's keys (``neck.0.*`` / ``neck.1.*``): xavier-uniform
    convolutions as the reference's init_cfg asks (segmentation/mmseg/models/necks/fpn.py:82-83), GroupNorm affine
    perturbed from (1, 0)."""
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


def make_bev_weights(feat_channels=512, num_layers=5, seed=0) -> Dict[str, torch.Tensor]:
    """BEV map-segmentation weights (bev/mmdet3d/models/fusion_models/ddp.py:65-116): the segmentation recipe with 6
    classes under the BEV tree's keys (``heads.map.*``) and a (256, feat_channels + 256) transform."""
    W = make_weights(task="seg", num_classes=6, num_layers=num_layers, seed=seed)
    if feat_channels != E:
        g = torch.Generator().manual_seed(seed + 7919)
        a = (6.0 / (feat_channels + 2 * E)) ** 0.5
        W["transform.conv.weight"] = (torch.rand(E, feat_channels + E, 1, 1, generator=g) * 2 - 1) * a
    return {("heads.map." + k[len("decode_head."):] if k.startswith("decode_head.") else k): v for k, v in W.items()}
