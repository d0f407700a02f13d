#!/bin/bash
# GPU calls 29 / 30 of round 2 (1 GPU): k_rank_sort over the touched word range with per-word prefixes (29), warp-per-source k_node_stats (30): makegraph parity on hardware
# (complete adjacency of every test plan incl. the full C2 graph), C5 bench subset.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bfs_schedules.py tests/test_gpu_fullsize.py -m gpu -x -q -p no:cacheprovider -k "makegraph or golden or row_ordering or c2_complete or capacity" > gpurun_out/r2c${CALL:-29}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2c${CALL:-29}_pytest.log
VGA_DEBUG_TIMING=1 timeout 400 python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r2c${CALL:-29}_default.json 2> gpurun_out/r2c${CALL:-29}_default.err
echo "== default rc=$?"
python - <<PY
import json
j = json.load(open("gpurun_out/r2c${CALL:-29}_default.json"))
s = j["stages"]
print("   value %.0f cells/s  build %.0f (sieve kernels %.0f, all kernels %.0f)  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s" % (
    j["value"], s["makegraph_ms"], s["sieve_kernels_ms"], s["makegraph_kernels_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
    j["result_checksum"]["sum_depth"]))
PY
