#!/bin/bash
# GPU call 4 of round 2 (1 GPU): bench C5 with the block cache + run-length local, local / sieve A/B, tensor-core probe,
# dram-bytes launch list, CLI -t rows.
mkdir -p gpurun_out
nproc > gpurun_out/r2c4_box.txt; free -g >> gpurun_out/r2c4_box.txt
VGA_BENCH_DEBUG=1 timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_C5.json 2> gpurun_out/r2c4_bench_C5.err
echo "bench C5 rc=$?"; tail -16 gpurun_out/r2c4_bench_C5.err; cut -c1-900 gpurun_out/r2c4_bench_C5.json
VGA_BENCH_DEBUG=1 timeout 600 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c4_bench_C2.json 2> gpurun_out/r2c4_bench_C2.err
echo "bench C2 rc=$?"; tail -4 gpurun_out/r2c4_bench_C2.err; cut -c1-600 gpurun_out/r2c4_bench_C2.json
{
  echo "== C5 build sieve_thread_cap=0"; VGA_DEBUG_TIMING=1 timeout 600 python tools/gpu_time.py C5 build
  echo "== C5 build sieve_thread_cap=1"; VGA_DEBUG_TIMING=1 timeout 600 python tools/gpu_time.py C5 build sieve_thread_cap=1
  echo "== C2 build sieve_thread_cap=0"; VGA_DEBUG_TIMING=1 timeout 600 python tools/gpu_time.py C2 build
  echo "== C2 build sieve_thread_cap=1"; VGA_DEBUG_TIMING=1 timeout 600 python tools/gpu_time.py C2 build sieve_thread_cap=1
  for P in C1 C2; do for M in 0 1 2; do
    echo "== $P local_mode=$M"; timeout 600 python tools/gpu_time.py $P local local_mode=$M
  done; done
  echo "== C4 local slice, modes 1 2"; timeout 900 python - <<'PY'
import time, numpy as np
from depthmapx_b200 import capi, plans
for name, lo, cnt in (("C4", 100000, 4096), ("C5", 500000, 4096)):
    flat = capi.prepare(plans.by_name(name))
    for mode in (1, 2):
        c = capi.Context(0); c.set_option("local_mode", mode); g = c.build(flat)
        g.local_ints((lo, lo + 64))
        t0 = time.time(); r = g.local_ints((lo, lo + cnt)); dt = time.time() - t0
        print(name, "local_mode", mode, cnt, "cells", round(dt, 3), "s", c.timing(), int(r[0].sum()), int(r[2].astype(np.int64).sum()), flush=True); c.close()
PY
} > gpurun_out/r2c4_ab.log 2>&1
timeout 600 python tools/local_tc_probe.py C1 > gpurun_out/r2c4_tc_probe_C1.json 2> gpurun_out/r2c4_tc_probe_C1.err; echo "tc probe C1 rc=$?"; cat gpurun_out/r2c4_tc_probe_C1.json
timeout 900 python tools/local_tc_probe.py C4 8192 > gpurun_out/r2c4_tc_probe_C4.json 2> gpurun_out/r2c4_tc_probe_C4.err; echo "tc probe C4 rc=$?"; cat gpurun_out/r2c4_tc_probe_C4.json; tail -3 gpurun_out/r2c4_tc_probe_C4.err
export VGA_TIME_SRC=8192
CMD="python tools/gpu_time.py C5 global"
$CMD > gpurun_out/r2c4_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c4_launches_dram_C5slice.csv $CMD > gpurun_out/r2c4_ncu_launches.log 2>&1
echo "dram launch list rc=$?"
unset VGA_TIME_SRC
timeout 1200 python tools/cli_timing.py C1 > gpurun_out/r2c4_cli_timing_C1.json 2> gpurun_out/r2c4_cli_timing.err; echo "cli timing rc=$?"; tail -3 gpurun_out/r2c4_cli_timing.err
ls -la gpurun_out | tail -12
