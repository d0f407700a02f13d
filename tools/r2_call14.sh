#!/bin/bash
# GPU call 14 of round 2 (1 GPU): faster y-major run lists (k_yruns: touched word range only, one warp per run, shuffle scans):
# parity of the hybrid lists, then the C5 / C2 bench lines with the lists' cost.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bfs_schedules.py tests/test_gpu_fullsize.py -m gpu -x -q -p no:cacheprovider -k "hybrid_x_or_y or (fullsize and not c4)" > gpurun_out/r2c14_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c14_pytest.log
VGA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --local-cells 0 > gpurun_out/r2c14_bench_default.json 2> gpurun_out/r2c14_bench_default.err
echo "bench rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c14_bench_default.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"build",round(d["stages"]["makegraph_ms"],1),"lists",round(d["stages"]["bfs_row_lists_ms"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),d["config"]["row_lists"], d["result_checksum"]["sum_depth"])
PY
timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c14_benchC2_default.json 2> gpurun_out/r2c14_benchC2_default.err
echo "bench C2 rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c14_benchC2_default.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"lists",round(d["stages"]["bfs_row_lists_ms"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),d["config"]["row_lists"], d["result_checksum"]["sum_depth"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c14_launches_lists.csv -k regex:"k_yruns|k_copy_chosen|k_trans|k_emit_nodes|k_run_costs|k_rows_copy|k_count_runs|k_mark_runs|k_emit_runs|k_choose" env VGA_TIME_SRC=1024 python tools/gpu_time.py C5 global > gpurun_out/r2c14_ncu.log 2>&1
echo "ncu rc=$?"
