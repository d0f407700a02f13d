#!/bin/bash
# ncu evidence (run under gpurun): launch list + full captures of the heavy BFS launches.
mkdir -p gpurun_out
CMD="python tools/gpu_time.py C2 global"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# heavy launches: k_push #2 (level 1->2), k_pull #2,#3 (levels 2->3, 3->4) of the first global run
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_push' -s 1 -c 2 -o gpurun_out/prof_push_r1b $CMD > gpurun_out/ncu_push.log 2>&1
echo "push capture rc=$?"
$CMD > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_pull' -s 1 -c 3 -o gpurun_out/prof_pull_r1b $CMD > gpurun_out/ncu_pull.log 2>&1
echo "pull capture rc=$?"
CMD2="python tools/gpu_time.py C2 build"
$CMD2 > gpurun_out/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_sieve_thread' -c 2 -o gpurun_out/prof_sievet_r1b $CMD2 > gpurun_out/ncu_sievet.log 2>&1
echo "sieve capture rc=$?"
ls -la gpurun_out | tail -12
