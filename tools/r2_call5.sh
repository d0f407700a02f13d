#!/bin/bash
# GPU call 5 of round 2 (1 GPU): word-width / lane-cooperative A/B on the bench subset itself, tensor-core probe,
# compute-sanitizer memcheck over small parity cases.
mkdir -p gpurun_out
for O in "" "--opt bfs_words=2" "--opt bfs_words=4 --opt bfs_coop=1" "--opt bfs_words=8 --opt bfs_coop=1" "--opt bfs_words=2 --opt sieve_thread_cap=0"; do
  T=$(echo "$O" | tr -d ' -' | tr '=' '_'); T=${T:-default}
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --local-cells 0 $O > gpurun_out/r2c5_bench_$T.json 2> gpurun_out/r2c5_bench_$T.err
  echo "bench [$O] rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c5_bench_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"build",round(d["stages"]["makegraph_ms"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),"batch",d["config"].get("bfs_batch_sources"),"frac",round(d["roofline"]["frac"],3), d["result_checksum"]["sum_depth"])
PY
done
for O in "" "--opt bfs_words=4 --opt bfs_coop=1" "--opt bfs_words=4"; do
  T=$(echo "$O" | tr -d ' -' | tr '=' '_'); T=${T:-default}
  timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $O > gpurun_out/r2c5_benchC2_$T.json 2> gpurun_out/r2c5_benchC2_$T.err
  echo "bench C2 [$O] rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c5_benchC2_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"bfs",round(d["stages"]["global_bfs_ms"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),"batch",d["config"].get("bfs_batch_sources"))
PY
done
timeout 600 python tools/local_tc_probe.py C1 > gpurun_out/r2c5_tc_probe_C1.json 2> gpurun_out/r2c5_tc_probe_C1.err; echo "tc probe C1 rc=$?"; cat gpurun_out/r2c5_tc_probe_C1.json; tail -2 gpurun_out/r2c5_tc_probe_C1.err
timeout 900 python tools/local_tc_probe.py C4 4096 65536 > gpurun_out/r2c5_tc_probe_C4.json 2> gpurun_out/r2c5_tc_probe_C4.err; echo "tc probe C4 rc=$?"; cat gpurun_out/r2c5_tc_probe_C4.json; tail -2 gpurun_out/r2c5_tc_probe_C4.err
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 python -m pytest -x -q -p no:cacheprovider -m gpu \
  tests/test_gpu_parity.py tests/test_gpu_bfs_schedules.py tests/test_stepdepth_gpu.py \
  -k "(golden and box2x2) or (golden and oblique20) or ghosts or empty or several_passes or (schedules and oblique and not 8-2) or (cooperative and oblique and 4) or (explicit and oblique) or round_trip or (row_ordering and holes) or step_depth" \
  > gpurun_out/r2c5_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r2c5_memcheck.log
