#!/bin/bash
# ncu evidence for round 1 (run under gpurun): launch list + full captures of the top kernels.
mkdir -p gpurun_out
CMD="python tools/gpu_time.py C2 build,global"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_push|k_pull|k_update' -s 14 -c 6 -o gpurun_out/prof_bfs_r1 $CMD > gpurun_out/ncu_bfs.log 2>&1
echo "bfs capture rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sieve' -c 2 -o gpurun_out/prof_sieve_r1 $CMD > gpurun_out/ncu_sieve.log 2>&1
echo "sieve capture rc=$?"
tail -3 gpurun_out/ncu_bfs.log gpurun_out/ncu_sieve.log
ls -la gpurun_out
