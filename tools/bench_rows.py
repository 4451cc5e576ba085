#!/usr/bin/env python
"""Device timing of the "next" rows of the scope table that sit beside the decode loop:

    python tools/bench_rows.py neck [--images 8] [--h 128 --w 256] [--channels 96 192 384 768] [--steps 10]
    python tools/bench_rows.py bev  [--images 1] [--randsteps 5] [--timesteps 3] [--feat 256] [--steps 5]
    python tools/bench_rows.py latency [--timesteps 10] [--steps 20]     # one image per call, with / without DDP_B200_GRAPH

Same timing rules as bench.py (>= 3 warm-up calls, CUDA events on the launching stream, synchronise on both sides,
inputs resident in HBM and larger than L2 or an explicit L2 flush inside the timed region).  One JSON line per run.
bench.py (the driver's contract: the decode loop's images/s) is deliberately left untouched by these rows.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps, flush):
    for _ in range(3):
        if flush is not None:
            flush.zero_()        # the fill kernel's first use (lazy module load) belongs to the warm-up, not the timed region
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def neck_flops(chans, h, w):
    """Algorithmic multiply-adds x 2 per image: laterals, 3x3 convs, per-level `down` slices."""
    total, hh, ww = 0, h, w
    for c in chans:
        n = hh * ww
        total += n * (2 * c * 256 + 2 * 9 * 256 * 256 + 2 * 256 * 256)
        hh, ww = (hh + 1) // 2, (ww + 1) // 2
    return total


def run_neck(a):
    from ddp_b200 import NeckEngine, synthetic
    g = torch.Generator().manual_seed(1)
    W = synthetic.make_neck_weights(a.channels, seed=7)
    eng = NeckEngine(a.channels)
    eng.load_state_dict(W)
    xs, hh, ww = [], a.h, a.w
    for c in a.channels:
        xs.append(torch.randn(a.images, c, hh, ww, generator=g).cuda())
        hh, ww = (hh + 1) // 2, (ww + 1) // 2
    in_bytes = sum(x.numel() * 4 for x in xs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if in_bytes < (192 << 20) else None
    ms = timed(lambda: eng.forward(xs), a.steps, flush)
    # per-call device times (events around single calls) next to the back-to-back average above
    per_call = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        eng.forward(xs)
        e1.record()
        torch.cuda.synchronize()
        per_call.append(round(e0.elapsed_time(e1), 3))
    fl = neck_flops(a.channels, a.h, a.w) * a.images
    print(json.dumps({"row": "neck (FPN + MultiStageMerging)", "images": a.images, "tokens": [a.h, a.w],
                      "in_channels": a.channels, "ms_per_call": ms, "images_per_s": a.images / (ms / 1e3),
                      "algorithmic_tflops": fl / (ms / 1e3) / 1e12, "launches": eng.last_launch_count,
                      "single_call_ms": per_call,
                      "arithmetic": "3x3 convs: tc_3xf16 tcgen05 implicit GEMM; 1x1 convs: fp32 CUDA-core GEMM"
                      if os.environ.get("DDP_B200_NECK_TC", "1") != "0" else "fp32 CUDA-core GEMMs (first path)",
                      "l2": "flushed before every call" if flush is not None else "inputs larger than L2"}))


def run_bev(a):
    from ddp_b200.bev import BevDecodeEngine, grid_coords
    from ddp_b200 import synthetic
    scopes = (((-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8)), ((-50.0, 50.0, 0.5), (-50.0, 50.0, 0.5)))    # shipped grid_transform
    gy, gx = grid_coords(*scopes)
    W = synthetic.make_bev_weights(feat_channels=a.feat, num_layers=5, seed=7)
    eng = BevDecodeEngine(timesteps=a.timesteps, feat_channels=a.feat, num_layers=5, gemm_mode=a.gemm)
    eng.load_state_dict(W)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(a.images, a.feat, 128, 128, generator=g).cuda()
    noise = torch.randn(a.images, a.randsteps, 256, 128, 128, generator=g).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: eng.sample(x, noise, gy, gx), a.steps, flush)
    per_layer = 2 * 256 * (256 + 64 + 32 + 256) + 2 * 2 * 256 * 1024
    fl = a.images * a.randsteps * a.timesteps * (200 * 200 * (5 * per_layer + 2 * 256 * 6) + 128 * 128 * 2 * 256 * 256)
    print(json.dumps({"row": "BEV map segmentation loop", "images": a.images, "randsteps": a.randsteps,
                      "timesteps": a.timesteps, "state_grid": [128, 128], "map_grid": [200, 200], "feat_channels": a.feat,
                      "ms_per_call": ms, "samples_per_s": a.images / (ms / 1e3),
                      "algorithmic_tflops": fl / (ms / 1e3) / 1e12, "launches": eng.last_launch_count, "gemm_mode": a.gemm,
                      "l2": "flushed before every call"}))


def run_latency(a):
    """The reference's own operating point: one image per GPU (tools/test.py, samples_per_gpu=1).  Per-call latency of the
    decode loop at the headline geometry with ordinary launches and with the CUDA-graph replay (DDP_B200_GRAPH=1)."""
    from ddp_b200 import DecodeEngine, synthetic
    W = synthetic.make_weights(task="seg", num_classes=19, seed=7)
    x, noise = synthetic.make_inputs("seg", 1, 1, 128, 256, seed=3)
    x, noise = x.cuda(), noise.cuda()
    out = torch.empty(1, 19, 128, 256, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    side = torch.cuda.Stream()
    for name, env in (("launches", None), ("graph", "1")):
        if env is None:
            os.environ.pop("DDP_B200_GRAPH", None)
        else:
            os.environ["DDP_B200_GRAPH"] = env
        eng = DecodeEngine(task="seg", num_classes=19, timesteps=a.timesteps, gemm_mode="tc_3xf16")
        eng.load_state_dict(W)
        torch.cuda.synchronize()
        with torch.cuda.stream(side):                       # graph capture needs a non-default stream
            res[name] = timed(lambda: eng.sample(x, noise, out=out), a.steps, flush)
        res[name + "_launches_per_call"] = eng.last_launch_count
    print(json.dumps({"row": "decode loop, one image per call (latency mode)", "tokens": [128, 256], "timesteps": a.timesteps,
                      "ms_per_image_launches": res["launches"], "ms_per_image_graph": res["graph"],
                      "launches_per_call": res["launches_launches_per_call"], "l2": "flushed before every call"}))


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="row", required=True)
    n = sub.add_parser("neck")
    n.add_argument("--images", type=int, default=8)
    n.add_argument("--h", type=int, default=128)
    n.add_argument("--w", type=int, default=256)
    n.add_argument("--channels", type=int, nargs="+", default=[96, 192, 384, 768])
    n.add_argument("--steps", type=int, default=10)
    b = sub.add_parser("bev")
    b.add_argument("--images", type=int, default=1)
    b.add_argument("--randsteps", type=int, default=5)
    b.add_argument("--timesteps", type=int, default=3)
    b.add_argument("--feat", type=int, default=256)
    b.add_argument("--gemm", default="tc_3xf16", choices=["fp32", "tc_3xf16", "tc_f16"])
    b.add_argument("--steps", type=int, default=20)
    lt = sub.add_parser("latency")
    lt.add_argument("--timesteps", type=int, default=10)
    lt.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    {"neck": run_neck, "bev": run_bev, "latency": run_latency}[a.row](a)


if __name__ == "__main__":
    main()
