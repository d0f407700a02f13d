#!/bin/bash
# Second GPU call of the next round (after tools/r2_first_call.sh picked the configuration): ncu evidence for it.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r2_profile.sh "bfs_push=1 bfs_pull=1 bfs_coarse=0"'
# Launch list of the whole step, then full captures of the heaviest launches of each pyramid kernel and of k_update;
# copy the summaries (python tools/ncu_summary.py <rep>) into profiles/ with an r2 prefix.
OPTS="${1:-bfs_push=1 bfs_pull=1 bfs_coarse=0}"
mkdir -p gpurun_out
CMD="python tools/gpu_time.py C2 global $OPTS"
$CMD > gpurun_out/r2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in k_push_pyr k_pull_pyr k_pyr_down k_pyr_build k_update k_push; do
  $CMD > gpurun_out/r2_plain_$K.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"^.*${K}[<(]" -s 2 -c 3 -o gpurun_out/r2_prof_$K $CMD > gpurun_out/r2_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2_prof_$K.ncu-rep > gpurun_out/r2_prof_${K}_summary.txt 2>/dev/null
done
ls -la gpurun_out | tail -20
