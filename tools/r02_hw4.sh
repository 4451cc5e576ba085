#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw4
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n 8 "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run tests python -m pytest tests/test_gpu_parity.py -q -x -k "host or cfg5 or graph or batched" --durations=5
T=600 run bench python bench.py
T=300 run bench_chunks1 env DDP_B200_HOST_CHUNKS=1 python bench.py --no-also --no-cpu-baseline
T=300 run bench_chunks3 env DDP_B200_HOST_CHUNKS=3 python bench.py --no-also --no-cpu-baseline
T=300 run bench_chunks4 env DDP_B200_HOST_CHUNKS=4 python bench.py --no-also --no-cpu-baseline
echo done
