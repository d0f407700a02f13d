#!/bin/bash
# GPU call 17 of round 2 (1 GPU): the whole GPU suite on the final code (x/y-major lists, metric / angular VGA), the default
# bench line, DRAM launch list of the BFS with both list kinds, ncu captures of the new kernels.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 > gpurun_out/r2c17_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/r2c17_pytest.log
timeout 900 python bench.py > gpurun_out/r2c17_bench_default.json 2> gpurun_out/r2c17_bench_default.err
echo "bench default rc=$?"; tail -3 gpurun_out/r2c17_bench_default.err; cut -c1-600 gpurun_out/r2c17_bench_default.json
timeout 600 python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r2c17_bench_C2.json 2> gpurun_out/r2c17_bench_C2.err
echo "bench C2 rc=$?"; cut -c1-300 gpurun_out/r2c17_bench_C2.json
export VGA_TIME_SRC=8192 VGA_TIME_RADII=-1 VGA_TIME_REPS=1
CMD="python tools/gpu_time.py C5 global bfs_hybrid=2"
$CMD > gpurun_out/r2c17_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c17_launches_dram_C5slice_xy.csv $CMD > gpurun_out/r2c17_ncu_launches.log 2>&1
echo "dram launch list rc=$?"
for K in k_push_nodes_coop k_pyr_down; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/r2c17_prof_${K}_xy $CMD > gpurun_out/r2c17_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2c17_prof_${K}_xy.ncu-rep > gpurun_out/r2c17_prof_${K}_xy_summary.txt 2>/dev/null
done
unset VGA_TIME_SRC VGA_TIME_RADII VGA_TIME_REPS
CMD2="python tools/metric_time.py C2 4096 0"
ncu --set full --clock-control none --import-source on -k regex:k_metric_angular -c 1 -o gpurun_out/r2c17_prof_k_metric_angular $CMD2 > gpurun_out/r2c17_ncu_metric.log 2>&1
echo "k_metric_angular capture rc=$?"
python tools/ncu_summary.py gpurun_out/r2c17_prof_k_metric_angular.ncu-rep > gpurun_out/r2c17_prof_k_metric_angular_summary.txt 2>/dev/null
rm -f gpurun_out/r2c17_prof_k_pyr_down_xy.ncu-rep
ls -la gpurun_out | grep r2c17
