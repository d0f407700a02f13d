#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "global or ghost or boundary or golden" > gpurun_out/pytest_bfs2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_bfs2.log
for opts in "bfs_coarse=0 pull_alpha=1" "bfs_coarse=0 pull_alpha=1 pull_beta=2" "bfs_coarse=1 pull_alpha=4" "bfs_coarse=1 pull_alpha=1" "bfs_coarse=1 pull_alpha=2" "bfs_coarse=1 pull_alpha=8" "bfs_coarse=1 pull_alpha=4 bfs_group=4" "bfs_coarse=1 pull_alpha=4 bfs_group=1" "bfs_coarse=1 pull_alpha=16 bfs_group=1" "bfs_coarse=1 bfs_mode=1 bfs_group=1"; do
  echo "== $opts"
  timeout 300 python tools/gpu_time.py C2 global $opts 2>&1 | grep -E "rep1" | sed -e 's/h2d_ms.*main_kernel_ms/main_kernel_ms/'
done
