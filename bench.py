#!/usr/bin/env python
"""bench.py — images/sec of the DDP reverse-diffusion decode loop on B200 (see DESIGN.md section d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--gemm MODE]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU path (oracle port) on host cores

One "step" = one pass of the hot path (ddp_sample: T DDIM steps of the denoiser) over one batch of
synthetic images already resident in HBM.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> task, classes, (h, w) tokens, T, R, accumulation, images per GPU, BASELINE.json config
WORKLOADS = {
    # BASELINE.json configs[2] (the config the metric "512x1024, T=10" is quoted on): 64 images on 8 GPUs = 8 per GPU
    "cityscapes_512x1024_T10": dict(task="seg", C=19, h=128, w=256, T=10, R=1, acc=False, per_gpu=8, bit_scale=0.01),
    "cityscapes_512x1024_T3": dict(task="seg", C=19, h=128, w=256, T=3, R=1, acc=False, per_gpu=8, bit_scale=0.01),
    # configs[1]: Swin-T ADE20K 512x512, 150 classes, T=3 (accumulation=True in the ADE configs), batch 32 on 1 GPU
    "ade_512x512_T3": dict(task="seg", C=150, h=128, w=128, T=3, R=1, acc=True, per_gpu=32, bit_scale=0.01),
    # configs[3]: depth NYU 480x640, T=20, batch 32 on 4 GPUs = 8 per GPU
    "nyu_480x640_T20": dict(task="depth", C=1, h=120, w=160, T=20, R=1, acc=False, per_gpu=8, bit_scale=0.1),
    # configs[4]: uncertainty mode K=8 samples x T=10, batch 16 on 8 GPUs = 2 per GPU
    "cityscapes_uncertainty_K8_T10": dict(task="seg", C=19, h=128, w=256, T=10, R=8, acc=False, per_gpu=2, bit_scale=0.01),
    # configs[0]-sized smoke workload
    "plumbing_64x64_T1": dict(task="seg", C=19, h=64, w=64, T=1, R=1, acc=False, per_gpu=1, bit_scale=0.01),
}
E, FFN = 256, 1024


def flops_per_token_step(wl):
    """Algorithmic GEMM FLOPs (multiply-add = 2), SURVEY.md section 8d."""
    per_layer = 2 * E * (E + 64 + 32 + E) + 2 * 2 * E * FFN
    if wl["task"] == "seg":
        return 6 * per_layer + 2 * E * E + 2 * E * wl["C"]
    return 6 * per_layer + 2 * E + 2 * E * 9


def flops_per_image(wl):
    n = wl["h"] * wl["w"]
    return n * (wl["R"] * wl["T"] * flops_per_token_step(wl) + 2 * E * E)


# algorithmic flops / bytes per token of one launch of each kernel class (fp32 activations in HBM)
def class_cost(name, wl):
    C = wl["C"]
    table = {
        "value_proj": (2 * E * E, 4 * (E + E)),
        "sampling_proj": (2 * E * 96, 4 * (E + 96 + 96)),
        "qproj_fused": (2 * E * E + 2 * E * 96, 4 * (E + E + 96 + 96)),  # value + sampling from one read of q
        "msda_gather": (2 * 8 * 4 * 4 * 32 + 0, 4 * (E + 96 + E)),
        "out_proj_ln": (2 * E * E, 4 * (E + E + E)),
        "ffn1_gelu": (2 * E * FFN, 4 * (E + FFN)),
        "ffn2_ln_film": (2 * E * FFN, 4 * (FFN + E + E)),
        "ffn_fused": (2 * 2 * E * FFN, 4 * (E + E)),          # hidden activation never leaves the SM
        "head_in": (2 * E * E, 4 * (E + E + E)),
        "head_out": (2 * E * C, 4 * (E + C)),
        "cond": (2 * E * E, 4 * (E + E)),
        "step_update": (0, 4 * (C + 2 * E)),
    }
    return table.get(name, (0, 0))


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained"),
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return None
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def make_weights_and_inputs(wl, B, seed):
    """Synthetic tensors of the named shapes: seeded CPU generator."""
    from ddp_b200 import synthetic
    W = synthetic.make_weights(task=wl["task"], num_classes=wl["C"], seed=7)
    x, noise = synthetic.make_inputs(wl["task"], wl["R"], B, wl["h"], wl["w"], seed=seed)
    return W, x, noise


def oracle_cfg(wl):
    from oracle import ddp_oracle as O       # cpu_baseline / --impl reference legs only
    return O.OracleConfig(task=wl["task"], num_classes=wl["C"], timesteps=wl["T"], randsteps=wl["R"],
                          accumulation=wl["acc"], bit_scale=wl["bit_scale"])


def cpu_reference_step(cfg, W, x1, noise1, steps_sampled):
    """One bounded sample of the reference's CPU path: 1 image, `steps_sampled` DDIM steps."""
    from oracle import ddp_oracle as O
    import dataclasses
    c2 = dataclasses.replace(cfg, timesteps=steps_sampled)
    t0 = time.perf_counter()
    O.sample(W, c2, x1, noise1)
    return time.perf_counter() - t0


def pick_cpu_threads(cfg, W, x1, noise1):
    """The reference's CPU path does not scale to every core of a large host (measured on the 128-core GPU boxes:
    all 128 torch threads were ~10x slower per DDIM step than 8 threads on a small host).  The baseline should be the
    reference at its best, so time ONE DDIM step at a few thread counts and keep the fastest.  Returns (threads, log)."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, cores // 4, cores // 2, cores) if 1 <= c <= cores}) or [cores]
    torch.set_num_threads(cands[0])
    cpu_reference_step(cfg, W, x1, noise1, 1)                     # warm-up (first touch, thread pools)
    best, best_t, log = cands[0], float("inf"), {}
    for c in cands:
        torch.set_num_threads(c)
        t = cpu_reference_step(cfg, W, x1, noise1, 1)
        log[c] = round(t, 2)
        if t < best_t:
            best, best_t = c, t
        elif t > 1.5 * best_t:                                    # past the knee: more threads only get worse
            break
    torch.set_num_threads(best)
    return best, log


def run_reference(args, wl, name):
    """--impl reference: the reference's own PyTorch-CPU path (oracle port, pinned bit-for-bit to the
    reference) on the host cores.  Each step = 1 image x 1 DDIM step; images/sec = 1 / (T * t_step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W, x, noise = make_weights_and_inputs(wl, 1, seed=1234)
    cfg = oracle_cfg(wl)
    if args.reference_device == "cuda":
        return run_reference_on_gpu(args, wl, name, cfg, W, x, noise)
    cores, thread_log = pick_cpu_threads(cfg, W, x, noise)
    for _ in range(args.warmup):
        cpu_reference_step(cfg, W, x, noise, 1)
    ts = [cpu_reference_step(cfg, W, x, noise, 1) for _ in range(args.steps)]
    t_step = sum(ts) / len(ts)
    value = 1.0 / (wl["T"] * t_step)
    line = {
        "impl": "reference", "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, name, 1, args),
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"1 image x 1 of {wl['T']} DDIM steps per bench step, images/s = 1/(T*t_step); "
                                   f"torch threads chosen by timing one step each: {thread_log} s of {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_on_gpu(args, wl, name, cfg, W, x, noise):
    """--impl reference --reference-device cuda (opt-in, SURVEY 8d "oracle-on-GPU"): the reference's own PyTorch op
    sequence (the oracle port) executed eagerly on cuda:0 — what the reference does on a GPU box except that its
    deformable attention goes through the grid_sample branch of mmcv's Python fallback instead of mmcv-full's CUDA op
    (not installable here).  One image per step (the reference loop is defined for b = 1), all T DDIM steps."""
    from oracle import ddp_oracle as O
    dev = torch.device("cuda", 0)
    Wd = {k: v.to(dev) for k, v in W.items()}
    xd, nd = x.to(dev), noise.to(dev)
    for _ in range(max(args.warmup, 1)):
        O.sample(Wd, cfg, xd, nd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        O.sample(Wd, cfg, xd, nd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    value = 1.0 / (ms / 1e3)
    line = {
        "impl": "reference", "reference_device": "cuda", "metric": "images/sec", "value": value, "unit": "images/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, name, 1, args),
        "note": "oracle-on-GPU: the reference's PyTorch op sequence run eagerly on the same B200, 1 image per call, all "
                f"{wl['T']} DDIM steps; deformable attention through grid_sample (mmcv's Python fallback), not mmcv-full's CUDA op",
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(wl, name, global_batch, args):
    return {"workload": name, "task": wl["task"], "tokens": [wl["h"], wl["w"]], "classes": wl["C"],
            "ddim_steps": wl["T"], "randsteps": wl["R"], "accumulation": wl["acc"], "global_batch": global_batch,
            "images_per_gpu": wl["per_gpu"], "parallelism": f"dp{args.gpus} (images sharded, one NCCL gather of logits)",
            "gemm_mode": args.gemm,
            "l2": ("inputs larger than L2 (x alone is %.0f MB per rank)" if wl["per_gpu"] * E * wl["h"] * wl["w"] * 4 >= (192 << 20)
                   else "x is %.0f MB per rank (< L2): a 256 MB buffer is overwritten before every step inside the timed region")
                  % (wl["per_gpu"] * E * wl["h"] * wl["w"] * 4 / 1e6)}


def measure(name, wl, args, ctx, steps, warmup, with_clocks=False):
    """One workload on this rank's GPU: clean device-resident timing (profiling OFF), a second pass with per-launch events
    for the kernel breakdown / roofline, the end-to-end pass through the host-buffer entry point, and (rank 0, N = 1) the two
    reference legs: the reference's op sequence on the same GPU and on the host cores.  Returns the JSON fields."""
    dev, world, rank, dist = ctx["dev"], ctx["world"], ctx["rank"], ctx["dist"]
    from ddp_b200 import DecodeEngine
    B = wl["per_gpu"]
    W, x, noise = make_weights_and_inputs(wl, B, seed=1234 + rank)
    eng = DecodeEngine(task=wl["task"], num_classes=wl["C"], timesteps=wl["T"], accumulation=wl["acc"],
                       bit_scale=wl["bit_scale"], gemm_mode=args.gemm)
    eng.load_state_dict(W)
    xd, nd = x.to(dev), noise.to(dev)
    xh, nh = x.pin_memory(), noise.pin_memory()
    out_h = torch.empty((B, eng.num_classes, wl["h"], wl["w"]), dtype=torch.float32).pin_memory()
    gathered = local_out = None
    if world > 1:
        gathered = torch.empty((world * B, eng.num_classes, wl["h"], wl["w"]), dtype=torch.float32, device=dev)
        local_out = torch.empty((B, eng.num_classes, wl["h"], wl["w"]), dtype=torch.float32, device=dev)

    # timing rule: inputs larger than L2, or an explicit flush between iterations.  The default workload's x is 268 MB
    # per rank; for the small ones (uncertainty: 67 MB, plumbing: 4 MB) a 256 MB buffer is overwritten before every step
    # INSIDE the timed region (~0.05 ms per step).
    x_bytes = B * E * wl["h"] * wl["w"] * 4
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if x_bytes < (192 << 20) else None

    out_fixed = torch.empty((B, eng.num_classes, wl["h"], wl["w"]), dtype=torch.float32, device=dev)

    def step_device():
        if flush_buf is not None:
            flush_buf.zero_()
        out = eng.sample(xd, nd, out=out_fixed)          # fixed result buffer (what DDP_B200_GRAPH=1 keys its CUDA graph on)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)     # the single gather of final logits
        return out

    def step_e2e():
        # synchronous host-buffer call: H2D x + noise, loop, D2H logits all inside, pipelined in two groups of images; at
        # N > 1 the result also lands in `local_out`, the send buffer of the one gather
        if flush_buf is not None:
            flush_buf.zero_()
        eng.sample_host(xh, nh, out=out_h, out_device=local_out)
        if world > 1:
            dist.all_gather_into_tensor(gathered, local_out)
        return out_h

    out_h2 = [out_h, torch.empty_like(out_h).pin_memory()]
    local_out2 = [local_out, torch.empty_like(local_out) if local_out is not None else None]

    def run_e2e_stream(n):
        """n calls through the streaming host entry point (two in flight): every call uploads its x + noise from pinned
        host memory and downloads its logits; the copies of call i+1 / i-1 overlap the loop of call i."""
        prev = None
        for i in range(n):
            if flush_buf is not None:
                flush_buf.zero_()
            t = eng.submit_host(xh, nh, out_h2[i % 2], out_device=local_out2[i % 2])
            if world > 1:
                dist.all_gather_into_tensor(gathered, local_out2[i % 2])
            if prev is not None:
                eng.wait_host(prev)
            prev = t
        eng.wait_host(prev)

    def timed(fn, n, sampler=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.start()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
            ms = float(t.item())
        return ms, clocks

    # (1) the headline: W warm-up steps, then exactly K steps, profiling OFF, device time between two events, max over ranks
    for _ in range(warmup):
        step_device()
    ms, clocks = timed(step_device, steps, ClockSampler(ctx["local"]) if (rank == 0 and with_clocks) else None)
    launches = eng.last_launch_count
    ms_per_step = ms / steps
    value = world * B * steps / (ms / 1e3)

    # (2) second pass: every launch bracketed by CUDA events on the launching stream -> per-class times, roofline
    psteps = max(1, min(steps, 3))
    eng.profile(True)
    ms_prof, _ = timed(step_device, psteps)
    prof = eng.profile_collect()
    eng.profile(False)

    # (3) end to end through the host-buffer entry points (pinned host x + noise in, logits out, every step):
    #     `e2e` = the streaming form (ddp_sample_host_submit / _wait, two calls in flight), the sustained rate of a queue
    #     of batches; `single_call` = one synchronous ddp_sample_host at a time (its own upload cannot hide)
    step_e2e()
    e2e_steps = max(1, min(steps, 3))
    ms_e2e1, _ = timed(step_e2e, e2e_steps)
    run_e2e_stream(2)
    ms_e2e, _ = timed(lambda: run_e2e_stream(steps), 1)
    e2e_value = world * B * steps / (ms_e2e / 1e3)
    e2e1_value = world * B * e2e_steps / (ms_e2e1 / 1e3)
    h2d = (xh.numel() + nh.numel()) * 4
    d2h = out_h.numel() * 4
    if rank != 0:
        return None

    peaks = load_peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][0])          # dominant kernel class by device time
    dname, (dms, dcount) = dom
    tokens_per_launch = B * wl["R"] * wl["h"] * wl["w"]
    fl, by = class_cost(dname, wl)
    avg_s = dms / max(dcount, 1) / 1e3
    gemm_like = dname not in ("msda_gather", "step_update", "finalize", "layout")
    if gemm_like:
        achieved = fl * tokens_per_launch / avg_s / 1e12
        roof = {"bound": "tensor", "kernel": dname, "achieved": achieved, "peak": peaks["tensor_sustained"] or peaks["tensor"],
                "unit": "TFLOP/s", "frac": achieved / (peaks["tensor_sustained"] or peaks["tensor"]),
                "peak_source": peaks["source"] + " (bf16 dense, sustained)", "traffic": None}
    else:
        achieved = by * tokens_per_launch / avg_s / 1e9
        roof = {"bound": "hbm", "kernel": dname, "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s",
                "frac": achieved / peaks["hbm"], "peak_source": peaks["source"], "traffic": None}
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel class: ncu cannot run inside the bench, so this
    # is the committed `ncu --set full` capture of the same command (profiles/ncu_traffic.json names the commit)
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        t = tj.get(f"{name}/{args.gemm}", {}).get(dname)
        if t is not None and wl["per_gpu"] == WORKLOADS[name]["per_gpu"]:
            roof["traffic"] = t
            roof["traffic_source"] = "profiles/ncu_traffic.json: " + str(tj.get("_source", "ncu --set full capture"))
    if gemm_like and args.gemm == "tc_3xf16":
        # fp32-faithful arithmetic on fp16 tensor cores issues three MMAs per algorithmic product: `frac`
        # counts the algorithmic FLOPs once, so its ceiling is 1/3; this is the tensor pipe's share of ITS ceiling
        roof["issued_mma_frac"] = 3 * roof["frac"]
    roof["algorithmic_bytes"] = by * tokens_per_launch
    roof["avg_launch_ms"] = 1e3 * avg_s
    roof["launches_timed"] = dcount
    total_prof = sum(v[0] for v in prof.values())
    roof["share_of_step"] = dms / total_prof if total_prof else None
    whole = world * B * steps * flops_per_image(wl) / (ms / 1e3) / 1e12
    kernel_ms = {k: round(v[0] / psteps, 3) for k, v in prof.items() if v[1]}

    res = {
        "value": value, "unit": "images/s", "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
        "config": dict(workload_config(wl, name, world * B, args),
                       host_numa=("rank bound to its GPU's NUMA node (%d cpus)" % len(ctx["numa_cpus"])) if ctx.get("numa_cpus") else "unbound"),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / steps, "steps": steps,
                "api": "ddp_sample_host_submit/_wait (streaming, two calls in flight, pinned host buffers)",
                "single_call": {"value": e2e1_value, "ms_per_step": ms_e2e1 / e2e_steps,
                                "api": "ddp_sample_host (synchronous, batch pipelined in two groups)"}},
        "gpu_launches": launches * steps, "launches_per_step": launches,
        "roofline": roof, "whole_loop_tflops": whole,
        "whole_loop_frac_of_tensor_peak": whole / (peaks["tensor_sustained"] or peaks["tensor"]),
        "kernel_ms_per_step": kernel_ms, "profiled_pass_ms_per_step": ms_prof / psteps,
    }
    if world == 1 and not args.no_cpu_baseline:
        del eng, xd, nd, flush_buf
        torch.cuda.empty_cache()
        cfg = oracle_cfg(wl)
        # the north star's own comparator: the reference's PyTorch op sequence on this same B200
        res["gpu_reference"] = gpu_reference(cfg, wl, W, x[:1], noise[:1], dev)
        cores, thread_log = pick_cpu_threads(cfg, W, x[:1], noise[:1])
        nsteps = 2 if wl["h"] * wl["w"] * wl["R"] >= 16384 else min(wl["T"], 4)
        t = cpu_reference_step(cfg, W, x[:1], noise[:1], nsteps)
        cpu_value = 1.0 / (t / nsteps * wl["T"])
        res["cpu_baseline"] = {"value": cpu_value, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": f"oracle (PyTorch-CPU port pinned to the reference), 1 image, {nsteps} of {wl['T']} "
                                         f"DDIM steps in {t:.1f} s, scaled to T={wl['T']}; torch threads chosen by timing one step "
                                         f"each: {thread_log} s of {os.cpu_count()} cores"}
    return res


def gpu_reference(cfg, wl, W, x1, noise1, dev, steps=2):
    """The reference's own PyTorch op sequence (oracle port) run eagerly on the same GPU, one image per call (the reference
    loop is defined for b = 1), all T steps.  Its deformable attention goes through the grid_sample branch of mmcv's Python
    fallback: mmcv-full's CUDA op is not installable here."""
    from oracle import ddp_oracle as O
    Wd = {k: v.to(dev) for k, v in W.items()}
    xd, nd = x1.to(dev), noise1.to(dev)
    O.sample(Wd, cfg, xd, nd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        O.sample(Wd, cfg, xd, nd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": 1e3 / ms, "unit": "images/s", "ms_per_image": ms,
            "note": "oracle-on-GPU: the reference's PyTorch op sequence, eager, fp32, on the same B200, 1 image per call, all "
                    f"{wl['T']} steps; deformable attention via grid_sample (mmcv's Python fallback), not mmcv-full's CUDA op"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cityscapes_512x1024_T10", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gemm", default=os.environ.get("DDP_GEMM", "tc_3xf16"), choices=["fp32", "tc_3xf16", "tc_f16"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU and oracle-on-GPU reference legs")
    ap.add_argument("--no-also", action="store_true", help="skip the other BASELINE.json configs in the `also` block")
    ap.add_argument("--reference-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cpu (default, the measurement contract) or cuda (oracle-on-GPU comparison)")
    ap.add_argument("--per-gpu", type=int, default=0, help="override images per GPU")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.per_gpu:
        wl["per_gpu"] = args.per_gpu
    if args.impl == "reference":
        return run_reference(args, wl, args.workload)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    numa_cpus = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # one process per GPU: keep each rank (and the pinned buffers it allocates) on its GPU's NUMA node.  N = 1 is left
        # alone: the cpu_baseline leg of that run uses every host core.  DDP_BENCH_NUMA=0 switches it off.
        if os.environ.get("DDP_BENCH_NUMA", "1") != "0":
            from ddp_b200.dist import bind_near_gpu
            numa_cpus = bind_near_gpu(local)
        dist.init_process_group("nccl", device_id=dev)
    ctx = dict(dev=dev, world=world, rank=rank, local=local, dist=dist, numa_cpus=numa_cpus)
    # everything below runs on a dedicated (non-default) stream: the library is asynchronous on the caller's stream, and the
    # opt-in CUDA-graph mode (DDP_B200_GRAPH=1) cannot capture the legacy default stream
    torch.cuda.synchronize()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    warmup = max(args.warmup, 3)
    res = measure(args.workload, wl, args, ctx, args.steps, warmup, with_clocks=True)
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # the metric names T = 3 as well, and BASELINE.json lists three more configs: each measured the same way (clean value,
    # e2e through host buffers, dominant-kernel roofline, GPU and CPU reference legs) with a short run, at N = 1 only
    also = {}
    if args.workload == "cityscapes_512x1024_T10" and world == 1 and not args.no_also:
        for other in ("cityscapes_512x1024_T3", "ade_512x512_T3", "nyu_480x640_T20", "cityscapes_uncertainty_K8_T10"):
            torch.cuda.empty_cache()
            also[other] = measure(other, dict(WORKLOADS[other]), args, ctx, 3, 3)

    line = {
        "metric": "images/sec", "value": res["value"], "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.gemm == "fp32" else ("f32 via 3xfp16 tensor-core split" if args.gemm == "tc_3xf16" else "f16"),
        "data": "synthetic", "config": res["config"], "clocks": res["clocks"], "e2e": res["e2e"],
        "gpu_launches": res["gpu_launches"], "launches_per_step": res["launches_per_step"],
        "roofline": res["roofline"], "whole_loop_tflops": res["whole_loop_tflops"],
        "whole_loop_frac_of_tensor_peak": res["whole_loop_frac_of_tensor_peak"],
        "kernel_ms_per_step": res["kernel_ms_per_step"], "profiled_pass_ms_per_step": res["profiled_pass_ms_per_step"],
        "gpu_reference": res.get("gpu_reference"), "cpu_baseline": res.get("cpu_baseline"), "also": also,
    }
    print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
