#!/bin/bash
# GPU call 8 of round 2 (1 GPU): the two bench arms as the driver runs them (reference first), host-phase timing on C2,
# word width with the lane-cooperative kernels on a C4 slice, direction thresholds on the bench subset.
mkdir -p gpurun_out
( time timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c8_reference_arm.json 2> gpurun_out/r2c8_reference_arm.err ) 2> gpurun_out/r2c8_reference_arm.time
echo "reference arm rc=$?"; tail -4 gpurun_out/r2c8_reference_arm.err; cat gpurun_out/r2c8_reference_arm.time; cut -c1-700 gpurun_out/r2c8_reference_arm.json
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c8_bench_C5.json 2> gpurun_out/r2c8_bench_C5.err ) 2> gpurun_out/r2c8_bench_C5.time
echo "bench C5 rc=$?"; tail -6 gpurun_out/r2c8_bench_C5.err; cat gpurun_out/r2c8_bench_C5.time; cut -c1-300 gpurun_out/r2c8_bench_C5.json
VGA_DEBUG_TIMING=1 VGA_BENCH_DEBUG=1 timeout 600 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c8_bench_C2_debug.json 2> gpurun_out/r2c8_bench_C2_debug.err
echo "bench C2 debug rc=$?"; tail -12 gpurun_out/r2c8_bench_C2_debug.err
{
  for W in 4 8; do echo "== C4 slice coop words=$W"; VGA_TIME_SRC=16384 timeout 600 python tools/gpu_time.py C4 global bfs_words=$W; done
  echo "== C1 default"; timeout 300 python tools/gpu_time.py C1 global
} > gpurun_out/r2c8_ab.log 2>&1
for O in "--opt pull_beta=2" "--opt pull_alpha=2"; do
  T=$(echo "$O" | tr -d ' -' | tr '=' '_')
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --local-cells 0 $O > gpurun_out/r2c8_bench_$T.json 2> gpurun_out/r2c8_bench_$T.err
  echo "bench [$O] rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c8_bench_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),"batch",d["config"].get("bfs_batch_sources"))
PY
done
