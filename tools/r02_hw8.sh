#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw8
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-6} "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=900 run tests python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py::test_full_size_cfg5_uncertainty_K8_T10_against_oracle --durations=5
T=600 run bench python bench.py --no-also --no-cpu-baseline
echo done
