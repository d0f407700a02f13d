#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "global or ghost or boundary or golden or properties" > gpurun_out/pytest_bfs3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_bfs3.log
for opts in "bfs_words=1" "bfs_words=2" "bfs_words=4" "bfs_words=4 bfs_mode=0" "bfs_words=4 bfs_mode=1" "bfs_words=4 pull_alpha=2" "bfs_words=4 pull_alpha=1 pull_beta=2" "bfs_words=4 bfs_coarse=0" "bfs_words=2 bfs_mode=0"; do
  echo "== $opts"
  timeout 300 python tools/gpu_time.py C2 global $opts 2>&1 | grep -E "rep1" | sed -e 's/h2d_ms.*main_kernel_ms/main_kernel_ms/'
done
