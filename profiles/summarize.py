#!/usr/bin/env python
"""Turn the scratch captures in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.md> "<title>"
    python profiles/summarize.py kernels  <report.ncu-rep> <out.md> "<title>"
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sass__inst_executed_local_loads", "local (spill) loads"),
]


def launches(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("ddp::tc::", "").replace("ddp::", "")
        agg[name][0] += 1
        agg[name][1] += float(d["Metric Value"].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` launch list (cold-cache, serialised: compare SHARES, not absolutes)\n\n")
        f.write(f"total {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / tot:.1f}% |\n")


def kernels(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`, one launch per kernel class (workload "
                "cityscapes_512x1024_T10, 8 images/GPU, 262144 tokens per launch).\n\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(CUtensorMap")[0].replace("void ", "")
            f.write(f"## `{name}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value |\n|---|---|\n")
            for k, label in KEYS:
                if k in idx:
                    f.write(f"| {label} (`{k}`) | {r[idx[k]]} {units[idx[k]]} |\n")
            try:
                def gb(k):
                    v, u = float(r[idx[k]]), units[idx[k]]
                    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
                tb = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
                f.write(f"| **traffic = dram read + write per launch** | {tb / 1e6:.1f} MB |\n")
                traffic[name] = tb
            except Exception:
                pass
            f.write("\n")
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels}[sys.argv[1]](*sys.argv[2:5])
