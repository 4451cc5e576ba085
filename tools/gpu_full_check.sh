#!/usr/bin/env bash
# full GPU suite (with parity log) + the default bench run, as the driver runs them
set -u
OUT=gpurun_out/full_check
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-900}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-6} "$OUT/$name.log" | cut -c1-600; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=300 run smoke python __graft_entry__.py smoke
T=1500 run gpu_suite python -m pytest tests -m gpu -q -x --durations=8
T=900 run bench python bench.py
T=300 run bench_reference python bench.py --impl reference --steps 3 --warmup 1
T=300 run bench_latency_T10 python tools/bench_rows.py latency --timesteps 10
T=300 run bench_latency_T3 python tools/bench_rows.py latency --timesteps 3
T=300 run bench_bev python tools/bench_rows.py bev
echo done
