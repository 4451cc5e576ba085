#!/usr/bin/env bash
# First hardware run of what was built without a GPU at the end of round 1 (neck: DESIGN.md §9, BEV loop: §10, the
# micro-benchmarks and the oracle-on-GPU comparator of profiles/r02_plan.md).  Meant for ONE gpurun call:
#
#   gpurun --timeout 2400 -- 'bash tools/first_hw_run.sh'
#
# Cheap, high-information steps first; the long ones (whole suite, compute-sanitizer) last.  Every step has its own
# timeout and log under gpurun_out/first_hw_run/; a failing step does not stop the next.
set -u
OUT=gpurun_out/first_hw_run
mkdir -p "$OUT"
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n 6 "$OUT/$name.log"; }

python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }

# 1. the pending tests as ordinary tests (--runxfail), neck first
T=600 run pending_neck python -m pytest tests/test_zz_gpu_neck.py -q --runxfail -rA
T=600 run pending_bev python -m pytest tests/test_zzz_gpu_bev.py -q --runxfail -rA
T=600 run pending_graph python -m pytest tests/test_zzzz_gpu_graph.py -q --runxfail -rA

# 2. device timing of both rows
T=300 run bench_neck python tools/bench_rows.py neck
T=300 run bench_bev python tools/bench_rows.py bev
T=300 run bench_bev_fusion python tools/bench_rows.py bev --feat 512
T=300 run bench_latency_T10 python tools/bench_rows.py latency --timesteps 10
T=300 run bench_latency_T3 python tools/bench_rows.py latency --timesteps 3

# 3. micro-benchmarks behind the round-2 kernel plan (built here if the binaries did not travel)
for u in ubench_gather ubench_sw_a; do
    [ -x tools/$u ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/$u tools/$u.cu -lcuda > "$OUT/build_$u.log" 2>&1
done
T=300 run ubench_gather_s05 tools/ubench_gather 0.5
T=300 run ubench_gather_s20 tools/ubench_gather 2.0
T=120 run ubench_sw_a tools/ubench_sw_a

# 4. the north star's comparator: the reference's PyTorch op sequence on the SAME B200 (opt-in arm)
T=600 run oracle_on_gpu_T10 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1
T=600 run oracle_on_gpu_T3 python bench.py --impl reference --reference-device cuda --steps 3 --warmup 1 --workload cityscapes_512x1024_T3

# 5. ncu launch lists of both rows
T=600 run ncu_neck ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/ncu_neck_launches.csv" \
    python tools/bench_rows.py neck --steps 1
T=600 run ncu_bev ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/ncu_bev_launches.csv" \
    python tools/bench_rows.py bev --steps 1

# 6. the whole GPU suite (verified tests first, the pending ones last and non-strict: tests/conftest.py)
T=1500 run gpu_suite python -m pytest tests -m gpu -q -rxX

# 7. compute-sanitizer on small cases of both rows
T=900 run sanitizer_neck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_neck.py -q -x --runxfail \
    -k "swin_l or error_behaviour"
T=900 run sanitizer_bev compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zzz_gpu_bev.py -q -x --runxfail \
    -k "fusion"
echo "done; see $OUT/"
