#!/bin/bash
# GPU call 26 of round 2 (1 GPU): k_push_delta with a forward walk through the segment prefix instead of a binary search per
# element: parity on hardware, per-level times on the C5 slice, C5 bench subset.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bfs_schedules.py -m gpu -x -q -p no:cacheprovider -k "delta_push or hybrid_x or whole_map" > gpurun_out/r2c26_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c26_pytest.log
export VGA_TIME_SRC=16384 VGA_TIME_RADII=-1 VGA_TIME_REPS=1 VGA_LEVEL_TIMING=1
timeout 300 python tools/gpu_time.py C5 global bfs_hybrid=2 > gpurun_out/r2c26_lt_hybrid.log 2>&1
echo "== lt_hybrid rc=$?"; grep -E "^\[level|^global" gpurun_out/r2c26_lt_hybrid.log | cut -c1-150
unset VGA_TIME_SRC VGA_TIME_RADII VGA_TIME_REPS VGA_LEVEL_TIMING
timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2c26_default.json 2> gpurun_out/r2c26_default.err
echo "== default rc=$?"
python - <<PY
import json
j = json.load(open("gpurun_out/r2c26_default.json"))
s = j["stages"]
print("   value %.0f cells/s  build %.0f (sieve kernels %.0f)  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s" % (
    j["value"], s["makegraph_ms"], s["sieve_kernels_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
    j["result_checksum"]["sum_depth"]))
PY
