#!/bin/bash
# GPU call 27 of round 2 (2 GPUs): the default bench on two ranks with the final kernels (x-major lists only on a rank:
# N / 2 sources per call is below the y-major threshold... at 2 ranks it is not: n/2 >= n/4), checksum equal to 1 GPU.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c27_bench_2gpu.json 2> gpurun_out/r2c27_bench_2gpu.err
echo "bench 2 GPUs rc=$?"; tail -3 gpurun_out/r2c27_bench_2gpu.err | cut -c1-300
python - <<PY
import json
j = json.loads([l for l in open("gpurun_out/r2c27_bench_2gpu.json") if l.startswith("{")][-1])
s = j["stages"]
print("   value %.0f cells/s  step %.0f ms  build %.0f  exchange %.1f  lists %.0f  bfs %.0f  level kernels %.0f  local %.0f  checksum %s e2e %.0f" % (
    j["value"], j["ms_per_step"], s["makegraph_ms"], s["exchange_ms"], s["bfs_row_lists_ms"], s["global_bfs_ms"], s["bfs_level_kernels_ms"], s["local_ms"],
    j["result_checksum"]["sum_depth"], j["e2e"]["value"]))
PY
