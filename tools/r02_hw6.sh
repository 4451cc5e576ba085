#!/usr/bin/env bash
set -u
OUT=gpurun_out/r02_hw6
mkdir -p "$OUT"
export DDP_PARITY_LOG=$PWD/$OUT/parity_log.jsonl
run() { local name=$1; shift; echo "== $name: $*"; timeout "${T:-600}" "$@" > "$OUT/$name.log" 2>&1; echo "   exit $? (log: $OUT/$name.log)"; tail -n ${TAILN:-8} "$OUT/$name.log"; }
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
T=600 run tests python -m pytest tests/test_gpu_parity.py -q -x -k "golden or teacher or pair or unfused or cfg3" --durations=5
T=600 run ncu_neck ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file "$OUT/ncu_neck_launches.csv" python tools/bench_rows.py neck --steps 1
T=600 run bench python bench.py --no-also --no-cpu-baseline
echo done
