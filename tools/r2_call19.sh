#!/bin/bash
# GPU call 19 of round 2 (1 GPU): delta push (k_push_delta, bfs_delta = 1) -- parity on hardware (BFS schedule tests; the C5 result checksum of the bench line must stay 1,029,504,649,975;
# full-size oracle samples in the final call), then A/B on the C5 bench subset: delta off / on, direction rule, nodes in flight.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bfs_schedules.py -m gpu -x -q -p no:cacheprovider --durations=5 > gpurun_out/r2c19_pytest.log 2>&1
echo "pytest rc=$?"; tail -9 gpurun_out/r2c19_pytest.log
ab() {
  T=$1; shift
  timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/r2c19_$T.json 2> gpurun_out/r2c19_$T.err
  echo "== $T rc=$? $*"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2c19_$T.json"))
    s = j["stages"]
    print("   value %.0f cells/s  build %.0f  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s" % (
        j["value"], s["makegraph_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
        j["result_checksum"]["sum_depth"]))
except Exception as e:
    print("   no line:", e)
PY
}
ab delta_default
ab delta_off --opt bfs_delta=0
ab delta_a1 --opt pull_alpha=1
ab delta_a1b2 --opt pull_alpha=1 --opt pull_beta=2
ab delta_a1b4 --opt pull_alpha=1 --opt pull_beta=4
ab delta_push_only --opt bfs_mode=0
ab delta_unroll2 --opt bfs_push_unroll=2
ls -la gpurun_out | grep r2c19
