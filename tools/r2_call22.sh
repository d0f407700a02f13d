#!/bin/bash
# GPU call 22 of round 2 (1 GPU): bit-sliced column counts in k_update, lane-cooperative pyramid passes, delta push with one
# node per lane in flight: parity on hardware, per-level times on the C5 slice (new / old pyramid passes), C5 bench subset A/B.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfs_schedules.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r2c22_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c22_pytest.log
export VGA_TIME_SRC=16384 VGA_TIME_RADII=-1 VGA_TIME_REPS=1 VGA_LEVEL_TIMING=1
run() {
  T=$1; shift
  timeout 300 python tools/gpu_time.py C5 global bfs_hybrid=2 "$@" > gpurun_out/r2c22_$T.log 2>&1
  echo "== $T rc=$? $*"; grep -E "^\[level|^global" gpurun_out/r2c22_$T.log | cut -c1-150
}
run lt_hybrid
run lt_hybrid_pyr_old bfs_pyr_coop=0
unset VGA_TIME_SRC VGA_TIME_RADII VGA_TIME_REPS VGA_LEVEL_TIMING
ab() {
  T=$1; shift
  timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/r2c22_$T.json 2> gpurun_out/r2c22_$T.err
  echo "== $T rc=$? $*"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2c22_$T.json"))
    s = j["stages"]
    print("   value %.0f cells/s  build %.0f  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s" % (
        j["value"], s["makegraph_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
        j["result_checksum"]["sum_depth"]))
except Exception as e:
    print("   no line:", e)
PY
}
ab default
ab w8 --opt bfs_delta_weight=8
ab w5 --opt bfs_delta_weight=5
ab words8 --opt bfs_words=8
ab words2 --opt bfs_words=2
