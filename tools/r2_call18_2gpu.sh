#!/bin/bash
# GPU call 18 of round 2 (2 GPUs): the sharded path with the x/y-major lists: bit-exactness vs one GPU, then the C5 bench at N = 2.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py office:128:128:5 > gpurun_out/r2c18_check.log 2>&1
echo "multi-gpu check rc=$?"; tail -3 gpurun_out/r2c18_check.log
VGA_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c18_bench_C5_2gpu.json 2> gpurun_out/r2c18_bench_C5_2gpu.err
echo "bench 2 gpus rc=$?"; tail -3 gpurun_out/r2c18_bench_C5_2gpu.err | cut -c1-300; tail -1 gpurun_out/r2c18_bench_C5_2gpu.json | cut -c1-1500
