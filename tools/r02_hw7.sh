#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw7
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-8} "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run neck_tests python -m pytest tests/test_zz_gpu_neck.py -q -rA --tb=short
T=300 run bench_neck_tc python tools/bench_rows.py neck
T=300 run bench_neck_fp32 env DDP_B200_NECK_TC=0 python tools/bench_rows.py neck
T=600 run ncu_neck ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file "$OUT/ncu_neck_launches.csv" python tools/bench_rows.py neck --steps 1
T=600 run sanitizer_neck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_neck.py -q -x -k "swin_l or ragged"
echo done
