#!/bin/bash
# GPU call 20 of round 2 (1 GPU): per-level device time of every kernel group (VGA_LEVEL_TIMING) on a C5 slice of 16,384
# sources (64 batches) with both list kinds: top-down only with / without delta push, bottom-up only, hybrid.
mkdir -p gpurun_out
export VGA_TIME_SRC=16384 VGA_TIME_RADII=-1 VGA_TIME_REPS=1 VGA_LEVEL_TIMING=1
run() {
  T=$1; shift
  timeout 300 python tools/gpu_time.py C5 global bfs_hybrid=2 "$@" > gpurun_out/r2c20_$T.log 2>&1
  echo "== $T rc=$? $*"; grep -E "^\[level|^global" gpurun_out/r2c20_$T.log | cut -c1-330
}
run push_delta bfs_mode=0
run push_delta_u2 bfs_mode=0 bfs_push_unroll=2
run push_old bfs_mode=0 bfs_delta=0
run pull_only bfs_mode=1
run hybrid_delta
run hybrid_old bfs_delta=0
