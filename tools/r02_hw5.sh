#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw5
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-8} "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run tests python -m pytest tests/test_gpu_parity.py -q -x -k "host or uncertainty_maps or graph" --durations=5
TAILN=30 T=200 run ubench_gather_s05 tools/ubench_gather 0.5
TAILN=30 T=200 run ubench_gather_s20 tools/ubench_gather 2.0
T=600 run bench python bench.py --no-also
echo done
