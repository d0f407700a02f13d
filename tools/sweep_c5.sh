#!/bin/bash
export VGA_TIME_SRC=65536
for opts in "bfs_words=1" "bfs_words=4" "bfs_words=4 bfs_mode=0" "bfs_words=2"; do
  echo "== C5 $opts"
  timeout 500 python tools/gpu_time.py C5 global $opts 2>&1 | grep -E "r=-1 rep1" | sed -e 's/h2d_ms.*main_kernel_ms/main_kernel_ms/'
done
for opts in "bfs_words=1" "bfs_words=4"; do
  echo "== C4 $opts"
  timeout 300 python tools/gpu_time.py C4 global $opts 2>&1 | grep -E "r=-1 rep1" | sed -e 's/h2d_ms.*main_kernel_ms/main_kernel_ms/'
done
