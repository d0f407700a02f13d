#!/bin/bash
# GPU call 25 of round 2 (1 GPU): final evidence of the round -- DRAM launch list of the BFS with the new kernels (-> measured
# traffic per source), default bench line (C5, with the host-buffer arm and the CPU baseline), the whole GPU suite, ncu
# captures of the top kernels, C2 bench line.
mkdir -p gpurun_out
export VGA_TIME_SRC=8192 VGA_TIME_RADII=-1 VGA_TIME_REPS=1
CMD="python tools/gpu_time.py C5 global bfs_hybrid=2"
$CMD > gpurun_out/r2c25_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c25_launches_dram_C5slice_delta.csv $CMD > gpurun_out/r2c25_ncu_launches.log 2>&1
echo "dram launch list rc=$?"
python tools/bfs_traffic.py gpurun_out/r2c25_launches_dram_C5slice_delta.csv C5 8192 profiles/r2_bfs_traffic.json gpurun_out/r2c25_bfs_traffic.json | tail -12
unset VGA_TIME_SRC VGA_TIME_RADII VGA_TIME_REPS
timeout 900 python bench.py > gpurun_out/r2c25_bench_default.json 2> gpurun_out/r2c25_bench_default.err
echo "bench default rc=$?"; tail -2 gpurun_out/r2c25_bench_default.err; cut -c1-700 gpurun_out/r2c25_bench_default.json
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 > gpurun_out/r2c25_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2c25_pytest.log
export VGA_TIME_SRC=8192 VGA_TIME_RADII=-1 VGA_TIME_REPS=1
for K in k_push_delta k_update k_pyr_down; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/r2c25_prof_$K $CMD > gpurun_out/r2c25_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2c25_prof_$K.ncu-rep > gpurun_out/r2c25_prof_${K}_summary.txt 2>/dev/null
done
unset VGA_TIME_SRC VGA_TIME_RADII VGA_TIME_REPS
rm -f gpurun_out/r2c25_prof_k_update.ncu-rep gpurun_out/r2c25_prof_k_pyr_down.ncu-rep
timeout 300 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r2c25_bench_C2.json 2> gpurun_out/r2c25_bench_C2.err
echo "bench C2 rc=$?"; cut -c1-300 gpurun_out/r2c25_bench_C2.json
ls -la gpurun_out | grep r2c25
