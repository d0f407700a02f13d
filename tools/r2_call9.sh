#!/bin/bash
# GPU call 9 of round 2 (1 GPU): tcgen05 int8 bring-up test, pull unroll A/B, CLI timing with the direct-Bin shim.
mkdir -p gpurun_out
timeout 60 ./tools/tc/umma_i8_test > gpurun_out/r2c9_umma_test.log 2>&1; echo "umma test rc=$?"; cat gpurun_out/r2c9_umma_test.log
nvidia-smi --query-gpu=name,memory.used --format=csv
for O in "" "--opt bfs_pull_unroll=4"; do
  T=$(echo "$O" | tr -d ' -' | tr '=' '_'); T=${T:-default}
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --local-cells 0 $O > gpurun_out/r2c9_bench_$T.json 2> gpurun_out/r2c9_bench_$T.err
  echo "bench [$O] rc=$?"; python - <<PY
import json
d=json.load(open("gpurun_out/r2c9_bench_$T.json"))
print("   value",round(d["value"]),"step",round(d["ms_per_step"],1),"level kernels",round(d["stages"]["bfs_level_kernels_ms"],1),"batch",d["config"].get("bfs_batch_sources"), d["result_checksum"]["sum_depth"])
PY
done
timeout 1200 python tools/cli_timing.py C1 > gpurun_out/r2c9_cli_timing_C1.json 2> gpurun_out/r2c9_cli_timing.err; echo "cli timing rc=$?"; tail -3 gpurun_out/r2c9_cli_timing.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c9_cli_timing_C1.json"))
for r in d["plans"]["C1"]["rows"]: print(r)
print({k:v for k,v in d["plans"]["C1"].items() if "identical" in k})
PY
