#!/bin/bash
# GPU call 24 of round 2 (1 GPU): emit pass of the sieve in order of decreasing task size: makegraph parity on hardware, C5
# bench subset with / without.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "makegraph or golden or overflow or shard or capacity" > gpurun_out/r2c24_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c24_pytest.log
ab() {
  T=$1; shift
  timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 "$@" > gpurun_out/r2c24_$T.json 2> gpurun_out/r2c24_$T.err
  echo "== $T rc=$? $*"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/r2c24_$T.json"))
    s = j["stages"]
    print("   value %.0f cells/s  build %.0f (sieve kernels %.0f)  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s" % (
        j["value"], s["makegraph_ms"], s["sieve_kernels_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
        j["result_checksum"]["sum_depth"]))
except Exception as e:
    print("   no line:", e)
PY
}
ab default
ab nosort --opt sieve_sort_emit=0
ab C2 --workload C2
ab C2_nosort --workload C2 --opt sieve_sort_emit=0
