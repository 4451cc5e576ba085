#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw16
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-900}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-4} "$OUT/$name.log" | cut -c1-300; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run tests python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_bev.py -q -x -k "golden or teacher or pair or cfg3 or host or bev or end_to_end" --tb=short
T=600 run bench python bench.py --no-also --no-cpu-baseline
T=600 run bench2 python bench.py --no-also --no-cpu-baseline
T=600 run sanitizer compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "seg_city_T3_R3_ragged or pair_options or host_pipeline"
echo done
