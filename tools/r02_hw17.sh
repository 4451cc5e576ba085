#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw17
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-900}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-4} "$OUT/$name.log" | cut -c1-300; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=900 run tests python -m pytest tests/test_gpu_parity.py -q -x --deselect tests/test_gpu_parity.py::test_full_size_cfg5_uncertainty_K8_T10_against_oracle --tb=short
T=600 run bench python bench.py --no-also --no-cpu-baseline
T=600 run bench_off env DDP_B200_COND_TC=0 python bench.py --no-also --no-cpu-baseline
T=600 run bench_T3 python bench.py --no-also --no-cpu-baseline --workload cityscapes_512x1024_T3
T=600 run bench_T3_off env DDP_B200_COND_TC=0 python bench.py --no-also --no-cpu-baseline --workload cityscapes_512x1024_T3
echo done
