#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw2
mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n 6 "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run neck python -m pytest tests/test_zz_gpu_neck.py -q -rA --tb=short
T=600 run bev python -m pytest tests/test_zzz_gpu_bev.py -q -rA --tb=short
T=300 run graph python -m pytest tests/test_zzzz_gpu_graph.py -q -rA --tb=short
T=300 run bench_neck python tools/bench_rows.py neck
T=300 run bench_bev python tools/bench_rows.py bev
T=300 run bench_latency_T10 python tools/bench_rows.py latency --timesteps 10
T=300 run bench_latency_T3 python tools/bench_rows.py latency --timesteps 3
T=900 run sanitizer_neck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_neck.py -q -x -k "swin_l or error_behaviour or ragged"
T=900 run sanitizer_bev compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zzz_gpu_bev.py -q -x -k "fusion"
echo done
