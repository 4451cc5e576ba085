#!/usr/bin/env bash
# ncu evidence of the default bench command: launch list (shares) + one --set full capture of the per-layer kernels.
set -u
TAG=${1:-r02}
OUT=gpurun_out/${TAG}_ncu
mkdir -p "$OUT"
python __graft_entry__.py > "$OUT/build.log" 2>&1 || { echo "build failed"; tail -n 30 "$OUT/build.log"; exit 1; }
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 822 -c 548 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > "$OUT/launches.log" 2>&1
echo "   exit $?"; tail -n 2 "$OUT/launches.log" | cut -c1-200
echo "== full capture"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ffn_fused|qproj_fused|msda_gather|gemm_tc_kernel|k_seg_step' -s 60 -c 10 \
    -o "$OUT/top" -f python bench.py --steps 1 --warmup 3 --no-also --no-cpu-baseline > "$OUT/full.log" 2>&1
echo "   exit $?"; tail -n 2 "$OUT/full.log" | cut -c1-200
ls -la "$OUT"
