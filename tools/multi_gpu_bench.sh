#!/usr/bin/env bash
# N-GPU bench exactly as the driver launches it
set -u
N=${1:-2}
OUT=gpurun_out/multi
mkdir -p "$OUT"
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_n$N.log" 2>&1
echo "exit $?"; grep '^{' "$OUT/bench_n$N.log" | tail -1 | cut -c1-1500; tail -3 "$OUT/bench_n$N.log" | cut -c1-300
