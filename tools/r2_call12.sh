#!/bin/bash
# GPU call 12 of round 2 (1 GPU): hybrid x-major / y-major lists: parity subset, then A/B on the bench workloads.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfs_schedules.py tests/test_gpu_fullsize.py -m gpu -x -q -p no:cacheprovider -k "not c4" > gpurun_out/r2c12_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c12_pytest.log
for O in "" "--opt bfs_hybrid=0"; do
  T=$(echo "$O" | tr -d ' -' | tr '=' '_'); T=${T:-default}
  VGA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --local-cells 0 $O > gpurun_out/r2c12_bench_$T.json 2> gpurun_out/r2c12_bench_$T.err
  echo "bench [$O] rc=$?"; grep "vga_global" gpurun_out/r2c12_bench_$T.err | tail -1; python - <<PY
import json
d=json.load(open("gpurun_out/r2c12_bench_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"build",round(d["stages"]["makegraph_ms"],1),"lists",round(d["stages"]["bfs_row_lists_ms"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),d["config"]["row_lists"], d["result_checksum"]["sum_depth"])
PY
  timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $O > gpurun_out/r2c12_benchC2_$T.json 2> gpurun_out/r2c12_benchC2_$T.err
  echo "bench C2 [$O] rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c12_benchC2_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"lists",round(d["stages"]["bfs_row_lists_ms"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),d["config"]["row_lists"], d["result_checksum"]["sum_depth"])
PY
done
