#!/usr/bin/env bash
# Round-2 first hardware call: tracebacks of the neck / BEV / graph tests, micro-benchmarks, GPU comparator, baseline bench.
set -u
OUT=gpurun_out/r02_hw1
mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n 4 "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=400 run pending_neck python -m pytest tests/test_zz_gpu_neck.py -q --runxfail -rA --tb=long -x -k "error_behaviour or ragged"
T=400 run pending_neck_all python -m pytest tests/test_zz_gpu_neck.py -q --runxfail -rA --tb=short
T=400 run pending_bev python -m pytest tests/test_zzz_gpu_bev.py -q --runxfail -rA --tb=long
T=300 run pending_graph python -m pytest tests/test_zzzz_gpu_graph.py -q --runxfail -rA
for u in ubench_gather ubench_sw_a; do
    [ -x tools/$u ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/$u tools/$u.cu -lcuda > "$OUT/build_$u.log" 2>&1
done
T=200 run ubench_gather_s05 tools/ubench_gather 0.5
T=200 run ubench_gather_s20 tools/ubench_gather 2.0
T=120 run ubench_sw_a tools/ubench_sw_a
T=400 run oracle_on_gpu_T10 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1
T=400 run oracle_on_gpu_T3 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1 --workload cityscapes_512x1024_T3
T=400 run bench_default python bench.py
echo done
