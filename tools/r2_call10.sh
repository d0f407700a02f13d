#!/bin/bash
# GPU call 10 of round 2 (1 GPU): the tensor-core local kernel against the oracle, then its timing on C1 / room plans.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "tensor_core" > gpurun_out/r2c10_pytest_tc.log 2>&1
echo "pytest tc rc=$?"; tail -15 gpurun_out/r2c10_pytest_tc.log
{
  for P in C1 room:60:60:2 C2; do for M in 1 4; do
    echo "== $P local_mode=$M"; timeout 300 python tools/gpu_time.py $P local local_mode=$M
  done; done
  echo "== C1 local auto"; timeout 300 python tools/gpu_time.py C1 local
} > gpurun_out/r2c10_ab.log 2>&1
grep -E "^==|^local" gpurun_out/r2c10_ab.log | cut -c1-230
