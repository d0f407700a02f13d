#!/bin/bash
mkdir -p gpurun_out
export VGA_TIME_SRC=8192
for H in 1 0; do
CMD="python tools/gpu_time.py C5 global bfs_hybrid=$H"
$CMD > gpurun_out/r2c13_plain_$H.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -c 4000 --csv --log-file gpurun_out/r2c13_launches_h$H.csv $CMD > gpurun_out/r2c13_ncu_$H.log 2>&1
echo "launch list hybrid=$H rc=$?"; grep "global r=-1 rep1" gpurun_out/r2c13_plain_$H.log | cut -c1-200
done
