#!/bin/bash
# GPU call 7 of round 2 (8 GPUs): the bench line on C5 at N = 8 and N = 4 (strong scaling of the fixed workload).
mkdir -p gpurun_out
for N in 8 4; do
  VGA_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2c7_bench_C5_${N}gpu.json 2> gpurun_out/r2c7_bench_C5_${N}gpu.err
  echo "bench $N gpus rc=$?"; grep -E "resident it=5" gpurun_out/r2c7_bench_C5_${N}gpu.err | sort | head -8; tail -1 gpurun_out/r2c7_bench_C5_${N}gpu.json | cut -c1-200
done
