"""Noise-schedule scalars of the sampling loop, evaluated on the host exactly as the reference does.

These are T x 5 scalars per call (not tensor work).  They are computed with the very torch ops, in
the very order, of the reference module-level functions so that the values handed to the CUDA
library are bit-identical to what the reference would use on CPU; log_snr(t) is ill-conditioned in
fp32 (cos near pi/2, 1/c^2 - 1 near t = 0), so "same formula, different rounding" is not enough.

Reference: segmentation/mmseg/models/segmentors/ddp.py:14-28, 204-213;
           depth/depth/models/depther/ddp.py:207-218.
"""
import math

import torch
from torch.special import expm1


def log(t, eps=1e-20):
    return torch.log(t.clamp(min=eps))


def beta_linear_log_snr(t):
    return -torch.log(expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t, ns=0.0002, ds=0.00025):
    return -log((torch.cos((t + ns) / (1 + ds) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


def log_snr_to_alpha_sigma(log_snr):
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


def gamma(t, ns=0.0002, ds=0.00025):
    return torch.cos(((t + ns) / (1 + ds)) * math.pi / 2) ** 2


def sampling_timesteps_seg(timesteps, time_difference, sample_range):
    times = []
    for step in range(timesteps):
        t_now = 1 - (step / timesteps) * (1 - sample_range[0])
        t_next = max(1 - (step + 1 + time_difference) / timesteps * (1 - sample_range[0]), sample_range[0])
        times.append((t_now, t_next))
    return times


def sampling_timesteps_depth(timesteps, time_difference):
    times = []
    for step in range(timesteps):
        t_now = 1 - step / timesteps
        t_next = max(1 - (step + 1 + time_difference) / timesteps, 0)
        times.append((t_now, t_next))
    return times


def seg_schedule(timesteps, time_difference, sample_range, noise_schedule):
    """-> five float lists (log_snr_now, alpha, sigma, alpha_next, sigma_next), one entry per step."""
    if noise_schedule == "linear":
        fn = beta_linear_log_snr
    elif noise_schedule == "cosine":
        fn = alpha_cosine_log_snr
    else:
        raise ValueError(f"invalid noise schedule {noise_schedule}")
    cols = [[], [], [], [], []]
    for t_now, t_next in sampling_timesteps_seg(timesteps, time_difference, sample_range):
        time = torch.tensor([t_now, t_next])
        l_now, l_next = fn(time[0:1]), fn(time[1:2])
        a, s = log_snr_to_alpha_sigma(l_now)
        an, sn = log_snr_to_alpha_sigma(l_next)
        for c, v in zip(cols, (l_now, a, s, an, sn)):
            c.append(float(v))
    return cols


def depth_schedule(timesteps, time_difference):
    """-> three float lists (t_now, gamma(t_now), gamma(t_next))."""
    cols = [[], [], []]
    for t_now, t_next in sampling_timesteps_depth(timesteps, time_difference):
        time = torch.tensor([t_now, t_next])
        for c, v in zip(cols, (time[0:1], gamma(time[0:1]), gamma(time[1:2]))):
            c.append(float(v))
    return cols


def seg_ddpm_schedule(timesteps, time_difference, sample_range, noise_schedule):
    """ddpm_sample scalars (ddp.py:274-284): -> (1 - c, c, exp(0.5 log variance), t_next > 0) per step."""
    fn = beta_linear_log_snr if noise_schedule == "linear" else alpha_cosine_log_snr
    cols = [[], [], [], []]
    for t_now, t_next in sampling_timesteps_seg(timesteps, time_difference, sample_range):
        time = torch.tensor([t_now, t_next])
        l_now, l_next = fn(time[0:1]), fn(time[1:2])
        _, sigma_next = log_snr_to_alpha_sigma(l_next)
        c = -expm1(l_now - l_next)
        variance = (sigma_next ** 2) * c
        std = (0.5 * log(variance)).exp()
        for col, v in zip(cols, (1 - c, c, std)):
            col.append(float(v))
        cols[3].append(int(bool(time[1] > 0)))
    return cols
