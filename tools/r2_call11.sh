#!/bin/bash
# GPU call 11 of round 2 (1 GPU): the whole GPU suite on the final code, the default bench line, ncu evidence of the
# final kernels (lane-cooperative BFS kernels on a C5 slice, tensor-core local kernel on C1).
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=8 > gpurun_out/r2c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/r2c11_pytest.log
timeout 900 python bench.py > gpurun_out/r2c11_bench_default.json 2> gpurun_out/r2c11_bench_default.err
echo "bench default rc=$?"; tail -3 gpurun_out/r2c11_bench_default.err; cut -c1-400 gpurun_out/r2c11_bench_default.json
export VGA_TIME_SRC=8192
CMD="python tools/gpu_time.py C5 global"
$CMD > gpurun_out/r2c11_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c11_launches_dram_C5slice.csv $CMD > gpurun_out/r2c11_ncu_launches.log 2>&1
echo "dram launch list rc=$?"
for K in k_push_nodes_coop k_pull_nodes_coop; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/r2c11_prof_$K $CMD > gpurun_out/r2c11_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2c11_prof_$K.ncu-rep > gpurun_out/r2c11_prof_${K}_summary.txt 2>/dev/null
done
unset VGA_TIME_SRC
CMD2="python tools/gpu_time.py C1 local"
$CMD2 > gpurun_out/r2c11_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_local_tc -c 1 -o gpurun_out/r2c11_prof_k_local_tc $CMD2 > gpurun_out/r2c11_ncu_k_local_tc.log 2>&1
echo "k_local_tc capture rc=$?"
python tools/ncu_summary.py gpurun_out/r2c11_prof_k_local_tc.ncu-rep > gpurun_out/r2c11_prof_k_local_tc_summary.txt 2>/dev/null
ls -la gpurun_out | grep r2c11
