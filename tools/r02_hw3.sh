#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw3
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n 8 "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=300 run smoke python __graft_entry__.py smoke
T=1500 run gpu_suite python -m pytest tests -m gpu -q -x --durations=12
echo done
