"""CPU oracle for the DDP reverse-diffusion decode loop.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain PyTorch-CPU ops and without any mmcv /
mmseg import, of the reference's "noise-to-map" sampling loop and of the
time-conditioned deformable-attention denoiser it calls.  It exists so that
the CUDA product under ``ddp_b200/`` can be checked; nothing in the product
path may import it (only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do).

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the *unmodified*
reference classes (``/root/reference/segmentation/mmseg/...`` with the
reference's own vendored mmcv 1.3.17 providing FFN / MultiScaleDeformable-
Attention / ConvModule) on seeded weights and inputs and stores inputs +
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
file against those fixtures (bit-for-bit at b=1).

Reference file:line followed by each function (paths relative to
/root/reference; "vmmcv" = controlnet/annotator/uniformer/mmcv):

  time_pairs_seg        segmentation/mmseg/models/segmentors/ddp.py:204-213
  log_snr_cosine/linear segmentation/mmseg/models/segmentors/ddp.py:14-24
  alpha_sigma           segmentation/mmseg/models/segmentors/ddp.py:27-28
  time_mlp              segmentation/mmseg/models/segmentors/ddp.py:31-46,103-112
  sine_pe               segmentation/mmseg/models/utils/transformer.py:78-113
  reference_points      segmentation/mmseg/models/decode_heads/deformable_head_with_time.py:63-88
  msda                  vmmcv/ops/multi_scale_deform_attn.py:94-151 (gather), 299-358 (forward)
  ffn                   vmmcv/cnn/bricks/transformer.py:253-279
  encoder_layer         segmentation/mmseg/models/utils/transformer.py:374-419
  head_tokens / head_seg segmentation/mmseg/models/decode_heads/deformable_head_with_time.py:90-132
  ddim_sample_seg       segmentation/mmseg/models/segmentors/ddp.py:215-246
  step_seg_one          segmentation/mmseg/models/segmentors/ddp.py:222-245 (one loop iteration; ddpm: 255-287)
  ddpm_sample_seg       segmentation/mmseg/models/segmentors/ddp.py:248-290
  time_pairs_depth      depth/depth/models/depther/ddp.py:210-218
  gamma_depth           depth/depth/models/depther/ddp.py:207-208
  head_depth            depth/depth/models/decode_heads/deformable_head_with_time.py:89-131,
                        depth/depth/models/decode_heads/decode_head.py:100,233-270
  sample_depth          depth/depth/models/depther/ddp.py:220-247, 97-110
  uncertainty           (no reference counterpart: defined here on the loop's per-step class maps, ddp.py:219,235,245)

Batched generalisation: the reference loop only works for one image (its
batch dimension is ``randsteps``); here rows are (image b, sample r) with
per-image means, and it reduces to the reference exactly at B=1.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

E = 256          # embed dims
HEADS = 8
POINTS = 4
FFN_DIM = 1024
LAYERS = 6
TIME_DIM = 1024


@dataclass
class OracleConfig:
    task: str = "seg"                 # "seg" | "depth"
    num_classes: int = 19
    timesteps: int = 3
    randsteps: int = 1
    time_difference: int = 1
    sample_range: Tuple[float, float] = (0.0, 0.999)
    noise_schedule: str = "cosine"
    bit_scale: float = 0.01
    accumulation: bool = False
    learned_sinusoidal_dim: int = 16
    num_layers: int = LAYERS
    min_depth: float = 1e-3
    max_depth: float = 10.0
    diffusion: str = "ddim"


# --------------------------------------------------------------------------
# synthetic weights: reference shapes + state-dict keys
# --------------------------------------------------------------------------
def make_weights(cfg: OracleConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
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
    cin = E if cfg.task == "seg" else 1
    if cfg.task == "seg":
        W["embedding_table.weight"] = randn(cfg.num_classes + 1, E)
        W["transform.conv.weight"] = xavier(E, 2 * E, 1, 1)
        W["transform.conv.bias"] = ubias(E)
    else:
        W["down.conv.weight"] = xavier(E, E + cin, 1, 1)
        W["down.conv.bias"] = ubias(E)
    W["time_mlp.0.weights"] = randn(cfg.learned_sinusoidal_dim // 2)
    W["time_mlp.1.weight"] = xavier(TIME_DIM, cfg.learned_sinusoidal_dim + 1)
    W["time_mlp.1.bias"] = ubias(TIME_DIM)
    W["time_mlp.3.weight"] = xavier(TIME_DIM, TIME_DIM)
    W["time_mlp.3.bias"] = ubias(TIME_DIM)
    thetas = torch.arange(HEADS, dtype=torch.float32) * (2.0 * math.pi / HEADS)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(HEADS, 1, 1, 2).repeat(1, 1, POINTS, 1)
    for i in range(POINTS):
        grid[:, :, i, :] *= i + 1
    for j in range(cfg.num_layers):
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
    if cfg.task == "seg":
        W["decode_head.conv_seg.weight"] = xavier(cfg.num_classes, E, 1, 1)
        W["decode_head.conv_seg.bias"] = ubias(cfg.num_classes)
    else:
        W["decode_head.conv_depth.weight"] = xavier(1, E, 3, 3)
        W["decode_head.conv_depth.bias"] = torch.full((1,), 3.0)
    return W


def cast_weights(W, dtype):
    return {k: v.to(dtype) for k, v in W.items()}


# --------------------------------------------------------------------------
# schedule
# --------------------------------------------------------------------------
def time_pairs_seg(cfg: OracleConfig) -> List[Tuple[float, float]]:
    T, s0, td = cfg.timesteps, cfg.sample_range[0], cfg.time_difference
    out = []
    for step in range(T):
        t_now = 1 - (step / T) * (1 - s0)
        t_next = max(1 - (step + 1 + td) / T * (1 - s0), s0)
        out.append((t_now, t_next))
    return out


def time_pairs_depth(cfg: OracleConfig) -> List[Tuple[float, float]]:
    T, td = cfg.timesteps, cfg.time_difference
    out = []
    for step in range(T):
        t_now = 1 - step / T
        t_next = max(1 - (step + 1 + td) / T, 0)
        out.append((t_now, t_next))
    return out


def _log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def log_snr_linear(t):
    return -torch.log(torch.special.expm1(1e-4 + 10 * (t ** 2)))


def log_snr_cosine(t, ns=0.0002, ds=0.00025):
    return -_log((torch.cos((t + ns) / (1 + ds) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def alpha_sigma(log_snr):
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


def gamma_depth(t, ns=0.0002, ds=0.00025):
    return torch.cos(((t + ns) / (1 + ds)) * math.pi / 2) ** 2


def time_mlp(W, l):
    """l: (b,) -> (b, 1024)."""
    x = l[:, None]
    freqs = x * W["time_mlp.0.weights"][None, :] * 2 * math.pi
    four = torch.cat((freqs.sin(), freqs.cos()), dim=-1)
    four = torch.cat((x, four), dim=-1)
    h = F.linear(four, W["time_mlp.1.weight"], W["time_mlp.1.bias"])
    h = F.gelu(h)
    return F.linear(h, W["time_mlp.3.weight"], W["time_mlp.3.bias"])


# --------------------------------------------------------------------------
# head
# --------------------------------------------------------------------------
def sine_pe(rows, h, w, dtype, num_feats=128, temperature=10000, scale=2 * math.pi,
            eps=1e-6, offset=-0.5, device=None):
    """(rows, 256, h, w); normalize=True as in every DDP config.  Built on `device` like the reference (the mask is
    allocated on feat.device, deformable_head_with_time.py:102)."""
    mask = torch.zeros((rows, h, w), device=device).to(torch.int)
    not_mask = 1 - mask
    y_embed = not_mask.cumsum(1, dtype=torch.float32)
    x_embed = not_mask.cumsum(2, dtype=torch.float32)
    y_embed = (y_embed + offset) / (y_embed[:, -1:, :] + eps) * scale
    x_embed = (x_embed + offset) / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32, device=device)
    dim_t = temperature ** (2 * (dim_t // 2) / num_feats)
    pos_x = x_embed[:, :, :, None] / dim_t
    pos_y = y_embed[:, :, :, None] / dim_t
    B, H, Wd = mask.size()
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).view(B, H, Wd, -1)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).view(B, H, Wd, -1)
    pos = torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)
    return pos.to(dtype)


def reference_points(h, w, dtype, device=None):
    ref_y, ref_x = torch.meshgrid(
        torch.linspace(0.5, h - 0.5, h, dtype=torch.float32, device=device),
        torch.linspace(0.5, w - 0.5, w, dtype=torch.float32, device=device), indexing="ij")
    ref_y = ref_y.reshape(-1)[None] / h
    ref_x = ref_x.reshape(-1)[None] / w
    ref = torch.stack((ref_x, ref_y), -1)
    return ref[:, :, None].to(dtype)            # (1, N, 1, 2)


def msda_gather(value, h, w, sampling_locations, attention_weights):
    """value (bs, N, 8, 32); loc (bs, Nq, 8, 1, 4, 2); aw (bs, Nq, 8, 1, 4)."""
    bs, _, num_heads, embed_dims = value.shape
    _, num_queries, _, num_levels, num_points, _ = sampling_locations.shape
    sampling_grids = 2 * sampling_locations - 1
    value_l_ = value.flatten(2).transpose(1, 2).reshape(bs * num_heads, embed_dims, h, w)
    sampling_grid_l_ = sampling_grids[:, :, :, 0].transpose(1, 2).flatten(0, 1)
    sampling_value_l_ = F.grid_sample(value_l_, sampling_grid_l_, mode="bilinear",
                                      padding_mode="zeros", align_corners=False)
    attention_weights = attention_weights.transpose(1, 2).reshape(
        bs * num_heads, 1, num_queries, num_levels * num_points)
    output = (torch.stack([sampling_value_l_], dim=-2).flatten(-2) * attention_weights) \
        .sum(-1).view(bs, num_heads * embed_dims, num_queries)
    return output.transpose(1, 2).contiguous()


def msda(W, p, query, query_pos, ref, h, w, taps=None):
    """query (N, bs, 256) -> (N, bs, 256)."""
    identity = query
    value = query
    query = query + query_pos
    query = query.permute(1, 0, 2)
    value = value.permute(1, 0, 2)
    bs, nq, _ = query.shape
    value = F.linear(value, W[p + "value_proj.weight"], W[p + "value_proj.bias"])
    value = value.view(bs, nq, HEADS, -1)
    off = F.linear(query, W[p + "sampling_offsets.weight"], W[p + "sampling_offsets.bias"]) \
        .view(bs, nq, HEADS, 1, POINTS, 2)
    aw = F.linear(query, W[p + "attention_weights.weight"], W[p + "attention_weights.bias"]) \
        .view(bs, nq, HEADS, POINTS)
    aw = aw.softmax(-1).view(bs, nq, HEADS, 1, POINTS)
    normalizer = torch.tensor([[w, h]], dtype=torch.long, device=query.device)
    loc = ref[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    out = msda_gather(value, h, w, loc, aw)
    if taps is not None:
        taps["value"] = value.reshape(bs, nq, E)
        taps["offsets"] = off.reshape(bs, nq, HEADS * POINTS * 2)
        taps["attn"] = aw.reshape(bs, nq, HEADS * POINTS)
        taps["gathered"] = out
    out = F.linear(out, W[p + "output_proj.weight"], W[p + "output_proj.bias"])
    out = out.permute(1, 0, 2)
    return out + identity


def encoder_layer(W, j, query, query_pos, ref, h, w, time, taps=None):
    p = f"decode_head.encoder.layers.{j}."
    query = msda(W, p + "attentions.0.", query, query_pos, ref, h, w, taps)
    query = F.layer_norm(query, (E,), W[p + "norms.0.weight"], W[p + "norms.0.bias"], 1e-5)
    if taps is not None:
        taps["ln1"] = query.permute(1, 0, 2)
    hid = F.gelu(F.linear(query, W[p + "ffns.0.layers.0.0.weight"], W[p + "ffns.0.layers.0.0.bias"]))
    out = F.linear(hid, W[p + "ffns.0.layers.1.weight"], W[p + "ffns.0.layers.1.bias"])
    query = query + out
    query = F.layer_norm(query, (E,), W[p + "norms.1.weight"], W[p + "norms.1.bias"], 1e-5)
    t = F.linear(F.silu(time), W[p + "time_mlp.1.weight"], W[p + "time_mlp.1.bias"])
    t = t[None]                                   # (1, b, 512)
    scale, shift = t.chunk(2, dim=2)
    if taps is not None:
        taps["film"] = t[0]
    query = query * (scale + 1) + shift
    return query


def head_tokens(W, cfg: OracleConfig, feat, time, taps: Optional[list] = None):
    """feat (rows, 256, h, w); time (1 or rows, 1024) -> memory (rows, 256, h, w)."""
    rows, c, h, w = feat.shape
    pos = sine_pe(rows, h, w, feat.dtype, device=feat.device).flatten(2).transpose(1, 2)
    q = feat.flatten(2).transpose(1, 2)
    ref = reference_points(h, w, feat.dtype, device=feat.device)
    q = q.permute(1, 0, 2)
    pos = pos.permute(1, 0, 2)
    for j in range(cfg.num_layers):
        lt = {} if taps is not None else None
        q = encoder_layer(W, j, q, pos, ref, h, w, time, lt)
        if taps is not None:
            lt["out"] = q.permute(1, 0, 2)
            taps.append(lt)
    memory = q.permute(1, 2, 0)
    return memory.reshape(rows, c, h, w).contiguous()


def head_seg(W, cfg, feat, time, taps=None):
    mem = head_tokens(W, cfg, feat, time, taps)
    return F.conv2d(mem, W["decode_head.conv_seg.weight"], W["decode_head.conv_seg.bias"])


def head_depth(W, cfg, feat, time, taps=None):
    mem = head_tokens(W, cfg, feat, time, taps)
    d = F.conv2d(mem, W["decode_head.conv_depth.weight"], W["decode_head.conv_depth.bias"], padding=1)
    return torch.relu(d) + cfg.min_depth


# --------------------------------------------------------------------------
# sampling loops
# --------------------------------------------------------------------------
@dataclass
class Trace:
    """Per-step intermediates (only filled when trace=True)."""
    feat: list = field(default_factory=list)       # (rows,256,h,w) head input
    temb: list = field(default_factory=list)       # (1,1024)
    layers: list = field(default_factory=list)     # per step: list of per-layer dicts
    logits: list = field(default_factory=list)     # (rows,C,h,w) or depth (rows,1,h,w)
    argmax: list = field(default_factory=list)
    mask_t: list = field(default_factory=list)     # state AFTER the update
    sched: list = field(default_factory=list)      # (log_snr, alpha, sigma, alpha_next, sigma_next)


def _rows(x, R):
    # einops repeat 'b c h w -> (r b) c h w' at b == 1 is x repeated R times
    return x.repeat(R, 1, 1, 1)


def step_seg_one(W, cfg: OracleConfig, xr, mask_t, idx, ddpm_noise_k=None, taps=None, sched_dtype=None):
    """ONE iteration of the reference loop body (ddp.py:222-245; ddpm: 255-287) for one image: xr (R,256,h,w) the
    repeated feature, mask_t (R,256,h,w) the state ENTERING step idx -> dict(logits, state, feat, temb, argmax, sched).
    `_sample_seg_one` is this function in a loop; the closed-loop parity check (tests/parity.py) calls it with the
    CUDA path's own state.  sched_dtype=torch.float32 with float64 tensors evaluates the schedule scalars (ill-conditioned
    in fp32) exactly as the fp32 reference does and everything downstream in float64 — the adjudicator."""
    dtype = xr.dtype
    sd = dtype if sched_dtype is None else sched_dtype
    log_snr_fn = log_snr_cosine if cfg.noise_schedule == "cosine" else log_snr_linear
    if cfg.noise_schedule not in ("cosine", "linear"):
        raise ValueError(f"invalid noise schedule {cfg.noise_schedule}")
    t_now, t_next = time_pairs_seg(cfg)[idx]
    times = torch.tensor([t_now, t_next], device=xr.device)          # float32, as in the reference
    times_now = times[0:1].to(sd)
    times_next = times[1:2].to(sd)
    feat = torch.cat([xr, mask_t], dim=1)
    feat = F.conv2d(feat, W["transform.conv.weight"], W["transform.conv.bias"])
    log_snr = log_snr_fn(times_now)
    log_snr_next = log_snr_fn(times_next)
    alpha, sigma = alpha_sigma(log_snr.view(1, 1, 1, 1))
    alpha_next, sigma_next = alpha_sigma(log_snr_next.view(1, 1, 1, 1))
    if sd != dtype:
        log_snr, log_snr_next = log_snr.to(dtype), log_snr_next.to(dtype)
        alpha, sigma, alpha_next, sigma_next = (v.to(dtype) for v in (alpha, sigma, alpha_next, sigma_next))
    temb = time_mlp(W, log_snr)
    mask_logit = head_seg(W, cfg, feat, temb, taps)
    pred_idx = torch.argmax(mask_logit, dim=1)
    mask_pred = F.embedding(pred_idx, W["embedding_table.weight"]).permute(0, 3, 1, 2)
    mask_pred = (torch.sigmoid(mask_pred) * 2 - 1) * cfg.bit_scale
    if cfg.diffusion == "ddim":
        pred_noise = (mask_t - alpha * mask_pred) / sigma.clamp(min=1e-8)
        mask_t = mask_pred * alpha_next + pred_noise * sigma_next
    elif cfg.diffusion == "ddpm":
        c = -torch.special.expm1(log_snr - log_snr_next)
        mean = alpha_next * (mask_t * (1 - c) / alpha + c * mask_pred)
        variance = (sigma_next ** 2) * c
        log_variance = _log(variance)
        nz = ddpm_noise_k if t_next > 0 else torch.zeros_like(mask_t)
        mask_t = mean + (0.5 * log_variance).exp() * nz
    else:
        raise NotImplementedError
    return dict(logits=mask_logit, state=mask_t, feat=feat, temb=temb, argmax=pred_idx,
                sched=tuple(float(v) for v in (log_snr, alpha, sigma, alpha_next, sigma_next)))


def _sample_seg_one(W, cfg: OracleConfig, x, noise, trace: Optional[Trace], ddpm_noise=None):
    """x (1,256,h,w); noise (R,256,h,w) -> (1,C,h,w).  Reference semantics (b=1)."""
    R = noise.shape[0]
    if cfg.noise_schedule not in ("cosine", "linear"):
        raise ValueError(f"invalid noise schedule {cfg.noise_schedule}")
    xr = _rows(x, R)
    mask_t = noise
    outs = []
    mask_logit = None
    for idx in range(len(time_pairs_seg(cfg))):
        taps = [] if trace is not None else None
        st = step_seg_one(W, cfg, xr, mask_t, idx, None if ddpm_noise is None else ddpm_noise[idx], taps)
        mask_logit, mask_t = st["logits"], st["state"]
        if cfg.accumulation:
            outs.append(mask_logit.softmax(1))
        if trace is not None:
            trace.feat.append(st["feat"])
            trace.temb.append(st["temb"])
            trace.layers.append(taps)
            trace.logits.append(mask_logit)
            trace.argmax.append(st["argmax"])
            trace.mask_t.append(mask_t)
            trace.sched.append(st["sched"])
    if cfg.accumulation:
        mask_logit = torch.cat(outs, dim=0)
    return mask_logit.mean(dim=0, keepdim=True)


def ddim_sample_seg(W, cfg: OracleConfig, x, noise, trace: bool = False, ddpm_noise=None):
    """x (B,256,h,w); noise (B,R,256,h,w) -> (B,C,h,w) [, list of per-image Trace]."""
    outs, traces = [], []
    with torch.no_grad():
        for b in range(x.shape[0]):
            tr = Trace() if trace else None
            outs.append(_sample_seg_one(W, cfg, x[b:b + 1], noise[b], tr,
                                        None if ddpm_noise is None else ddpm_noise[b]))
            traces.append(tr)
    out = torch.cat(outs, dim=0)
    return (out, traces) if trace else out


def _sample_depth_one(W, cfg: OracleConfig, x, noise, trace: Optional[Trace]):
    dtype = x.dtype
    R = noise.shape[0]
    xr = _rows(x, R)
    depth_t = noise
    depth_pred = None
    for (t_now, t_next) in time_pairs_depth(cfg):
        times = torch.tensor([t_now, t_next], device=x.device)
        times_now = times[0:1].to(dtype)
        times_next = times[1:2].to(dtype)
        feat = torch.cat([xr, depth_t], dim=1)
        feat = F.conv2d(feat, W["down.conv.weight"], W["down.conv.bias"])
        temb = time_mlp(W, times_now)
        taps = [] if trace is not None else None
        depth_pred = head_depth(W, cfg, feat, temb, taps)
        dn = (depth_pred - cfg.min_depth) / (cfg.max_depth - cfg.min_depth)
        dn = ((dn * 2) - 1) * cfg.bit_scale
        tn = times_now.view(1, 1, 1, 1)
        tx = times_next.view(1, 1, 1, 1)
        a_now = gamma_depth(tn)
        a_next = gamma_depth(tx)
        dn = dn.clamp(-cfg.bit_scale, cfg.bit_scale)
        eps = (1 / (1 - a_now).sqrt()) * (depth_t - a_now.sqrt() * dn)
        depth_t = a_next.sqrt() * dn + (1 - a_next).sqrt() * eps
        if trace is not None:
            trace.feat.append(feat)
            trace.temb.append(temb)
            trace.layers.append(taps)
            trace.logits.append(depth_pred)
            trace.mask_t.append(depth_t)
            trace.sched.append((float(a_now), float(a_next)))
    out = depth_pred.mean(dim=0, keepdim=True)
    return torch.clamp(out, min=cfg.min_depth, max=cfg.max_depth)


def sample_depth(W, cfg: OracleConfig, x, noise, trace: bool = False):
    """x (B,256,h,w); noise (B,R,1,h,w) -> (B,1,h,w), clamped as encode_decode does."""
    outs, traces = [], []
    with torch.no_grad():
        for b in range(x.shape[0]):
            tr = Trace() if trace else None
            outs.append(_sample_depth_one(W, cfg, x[b:b + 1], noise[b], tr))
            traces.append(tr)
    out = torch.cat(outs, dim=0)
    return (out, traces) if trace else out


def sample(W, cfg: OracleConfig, x, noise, trace: bool = False, ddpm_noise=None):
    """ddpm_noise (B, T, R, 256, h, w): the per-step noise of diffusion='ddpm'."""
    if cfg.task == "seg":
        return ddim_sample_seg(W, cfg, x, noise, trace, ddpm_noise)
    return sample_depth(W, cfg, x, noise, trace)


def uncertainty(W, cfg: OracleConfig, x, noise):
    """Per-pixel uncertainty maps of the library (ddp_set_uncertainty_outputs).  The reference has no such output: it
    exposes its stochastic samples only as ``randsteps`` + the mean (segmentation/mmseg/models/segmentors/ddp.py:219, 245;
    README abstract "uncertainty awareness"), so the definitions are made here, on the reference loop's own per-step
    class maps, and the CUDA path is checked against them:
      seg   changes (B,h,w) int32 = sum over samples r and steps k >= 1 of [argmax_k != argmax_{k-1}]
            spread  (B,h,w)       = 1 - mean_r [argmax of sample r at the LAST step == argmax of the returned map]
      depth spread  (B,h,w)       = population standard deviation over r of the last-step prediction
    -> (out, changes or None, spread)."""
    out, traces = sample(W, cfg, x, noise, trace=True)
    changes, spread = [], []
    for b, tr in enumerate(traces):
        if cfg.task == "seg":
            am = torch.stack(tr.argmax)                                   # (T,R,h,w)
            changes.append((am[1:] != am[:-1]).sum((0, 1)).to(torch.int32))
            spread.append(1.0 - (am[-1] == out[b].argmax(0)[None]).float().mean(0))
        else:
            spread.append(tr.logits[-1][:, 0].std(0, unbiased=False))
    return out, (torch.stack(changes) if changes else None), torch.stack(spread)


def make_inputs(cfg: OracleConfig, B, h, w, seed=1234, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    cin = E if cfg.task == "seg" else 1
    x = torch.randn(B, E, h, w, generator=g)
    noise = torch.randn(B, cfg.randsteps, cin, h, w, generator=g)
    return x.to(dtype), noise.to(dtype)
