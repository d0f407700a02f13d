#!/bin/bash
# GPU call 16 of round 2 (1 GPU): direction choice with the x/y-major lists (pull_alpha sweep on a C5 slice of 16,384 sources)
mkdir -p gpurun_out
export VGA_TIME_SRC=16384 VGA_TIME_RADII=-1 VGA_TIME_REPS=2
for A in "pull_alpha=1" "pull_alpha=2" "pull_alpha=3" "pull_alpha=5" "pull_alpha=2 pull_beta=3"; do
  T=$(echo "$A" | tr ' =' '__')
  timeout 300 python tools/gpu_time.py C5 global bfs_hybrid=2 $A > gpurun_out/r2c16_$T.log 2>&1
  echo "[$A] rc=$?"; grep "global r=-1 rep1" gpurun_out/r2c16_$T.log | sed 's/.*wall/wall/' | cut -c1-230
done
