#!/bin/bash
# GPU call 2 of round 2: ncu evidence for the node-list pyramid schedule on a C5 slice (the HBM-streaming case) + word-width /
# cost-weight sweep of the same schedule.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_profile_nodes.sh'
OPTS="bfs_push=1 bfs_pull=1 bfs_coarse=0 bfs_pyr_nodes=1"
mkdir -p gpurun_out
nproc > gpurun_out/r2_box.txt; free -g >> gpurun_out/r2_box.txt; nvidia-smi -L >> gpurun_out/r2_box.txt
{
  for W in 1 2 4; do
    echo "== C5 slice node lists words=$W"; VGA_TIME_SRC=32768 python tools/gpu_time.py C5 global $OPTS bfs_words=$W
  done
  for C in 50 200 400; do
    echo "== C5 slice node lists pyr_cost=$C"; VGA_TIME_SRC=32768 python tools/gpu_time.py C5 global $OPTS bfs_pyr_cost=$C
  done
  echo "== C5 slice node lists push only"; VGA_TIME_SRC=32768 python tools/gpu_time.py C5 global $OPTS bfs_mode=0
  echo "== C5 slice node lists pull only"; VGA_TIME_SRC=32768 python tools/gpu_time.py C5 global $OPTS bfs_mode=1
  echo "== C4 slice node lists words=2"; VGA_TIME_SRC=16384 python tools/gpu_time.py C4 global $OPTS bfs_words=2
  echo "== C1 node lists"; python tools/gpu_time.py C1 global $OPTS
  echo "== C1 default"; python tools/gpu_time.py C1 global
} > gpurun_out/r2_ab2.log 2>&1
export VGA_TIME_SRC=8192
CMD="python tools/gpu_time.py C5 global $OPTS"
$CMD > gpurun_out/r2n_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2n_launches_C5slice.csv $CMD > gpurun_out/r2n_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in k_push_nodes k_pull_nodes k_update k_pyr_build k_pyr_down; do
  ncu --set full --clock-control none --import-source on -k regex:"^.*${K}[<(]" -s 3 -c 3 -o gpurun_out/r2n_prof_$K $CMD > gpurun_out/r2n_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2n_prof_$K.ncu-rep > gpurun_out/r2n_prof_${K}_summary.txt 2>/dev/null
done
ls -la gpurun_out | tail -20
