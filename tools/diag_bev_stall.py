"""Diagnostic: per-call device time of the BEV loop, repeated, to localise an intermittent stall."""
import os, sys, json, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ddp_b200.bev import BevDecodeEngine, grid_coords
from ddp_b200 import synthetic

scopes = (((-51.2, 51.2, 0.8), (-51.2, 51.2, 0.8)), ((-50.0, 50.0, 0.5), (-50.0, 50.0, 0.5)))
gy, gx = grid_coords(*scopes)
W = synthetic.make_bev_weights(feat_channels=256, num_layers=5, seed=7)
eng = BevDecodeEngine(timesteps=3, feat_channels=256, num_layers=5, gemm_mode="tc_3xf16")
eng.load_state_dict(W)
g = torch.Generator().manual_seed(2)
x = torch.randn(1, 256, 128, 128, generator=g).cuda()
noise = torch.randn(1, 5, 256, 128, 128, generator=g).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
use_flush = "--no-flush" not in sys.argv
times, walls = [], []
for i in range(40):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    e0.record()
    if use_flush:
        flush.zero_()
    e1.record()
    eng.sample(x, noise, gy, gx)
    t1 = time.perf_counter()
    e2.record()
    torch.cuda.synchronize()
    times.append((round(e0.elapsed_time(e1), 2), round(e1.elapsed_time(e2), 2), round(1e3 * (t1 - t0), 2)))
print("flush_ms, sample_ms, host_enqueue_ms per call:")
print(times)

# back-to-back batches (no host synchronisation between the calls of a batch), as tools/bench_rows.py times them
batches = []
for b in range(30):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(5):
        flush.zero_()
        eng.sample(x, noise, gy, gx)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    batches.append((round(e0.elapsed_time(e1) / 5, 2), round(1e3 * (t1 - t0) / 5, 2)))
print("back-to-back batches of 5: (device ms per call, host enqueue ms per call):")
print(batches)
