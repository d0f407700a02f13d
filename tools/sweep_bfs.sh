#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "global or ghost or boundary or golden" > gpurun_out/pytest_bfs.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_bfs.log
for opts in "bfs_mode=0 bfs_order=0" "bfs_mode=0 bfs_order=1" "bfs_mode=2 bfs_order=0 bfs_coarse=0" "bfs_mode=2 bfs_order=1 bfs_coarse=0" "bfs_mode=2 bfs_order=1 bfs_coarse=1" "bfs_mode=2 bfs_order=1 bfs_coarse=1 pull_alpha=16" "bfs_mode=2 bfs_order=1 bfs_coarse=1 pull_alpha=64" "bfs_mode=1 bfs_order=1 bfs_coarse=1" "bfs_mode=2 bfs_order=1 bfs_coarse=1 bfs_group=4 pull_alpha=16" "bfs_mode=2 bfs_order=1 bfs_coarse=1 bfs_group=1 pull_alpha=16"; do
  echo "== $opts"
  timeout 300 python tools/gpu_time.py C2 global $opts 2>&1 | grep -E "global r" | sed -e 's/h2d_ms.*main_kernel_ms/main_kernel_ms/'
done
