#!/usr/bin/env bash
# N-GPU bench with and without the NUMA binding of bench.py, alternating on ONE box: prints value / e2e per run
set -u
N=${1:-2}; REPS=${2:-2}
OUT=gpurun_out/numa_ab
mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
for r in $(seq 1 $REPS); do
  for v in 1 0; do
    DDP_BENCH_NUMA=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/n${N}_numa${v}_r$r.log" 2>&1
    python - "$OUT/n${N}_numa${v}_r$r.log" $v <<'PY'
import json, sys
line = [l for l in open(sys.argv[1]) if l.startswith("{")]
if not line:
    print("numa", sys.argv[2], "no JSON line"); sys.exit(0)
d = json.loads(line[-1])
print("numa", sys.argv[2], "value %.2f e2e %.2f ratio %.4f single_call %.2f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"] / d["value"], d["e2e"]["single_call"]["value"]), d["config"].get("host_numa"))
PY
  done
done
