#!/bin/bash
# GPU call 6 of round 2 (2 GPUs): bit-exactness of the sharded path vs one GPU, then the bench line on C5 at N = 2.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py office:128:128:5 > gpurun_out/r2c6_check.log 2>&1
echo "multi-gpu check rc=$?"; tail -3 gpurun_out/r2c6_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/multi_gpu_check.py urban:160:160:4 > gpurun_out/r2c6_check2.log 2>&1
echo "multi-gpu check 2 rc=$?"; tail -3 gpurun_out/r2c6_check2.log
VGA_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c6_bench_C5_2gpu.json 2> gpurun_out/r2c6_bench_C5_2gpu.err
echo "bench 2 gpus rc=$?"; grep -E "rank 0|rank 1" gpurun_out/r2c6_bench_C5_2gpu.err | tail -8; tail -3 gpurun_out/r2c6_bench_C5_2gpu.err | cut -c1-300; tail -1 gpurun_out/r2c6_bench_C5_2gpu.json | cut -c1-1200
