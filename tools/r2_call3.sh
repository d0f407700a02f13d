#!/bin/bash
# GPU call 3 of round 2: the whole GPU suite on the rewritten BFS, first bench lines (small plan, then C5), word-width /
# direction-threshold sweeps of the node-list schedule, ncu launch list + full captures on a C5 slice.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=15 > gpurun_out/r2c3_pytest.log 2>&1
echo "pytest rc=$?"; tail -25 gpurun_out/r2c3_pytest.log
timeout 300 python bench.py --workload office:96:96:3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2c3_bench_small.json 2> gpurun_out/r2c3_bench_small.err
echo "bench small rc=$?"; tail -3 gpurun_out/r2c3_bench_small.err; cut -c1-600 gpurun_out/r2c3_bench_small.json
VGA_BENCH_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2c3_bench_C5.json 2> gpurun_out/r2c3_bench_C5.err
echo "bench C5 rc=$?"; tail -12 gpurun_out/r2c3_bench_C5.err; cut -c1-1500 gpurun_out/r2c3_bench_C5.json
{
  for W in 1 2 4 8; do
    echo "== C5 slice words=$W"; VGA_TIME_SRC=32768 timeout 600 python tools/gpu_time.py C5 global bfs_words=$W
  done
  for A in 2 4; do
    echo "== C5 slice words=2 pull_alpha=$A"; VGA_TIME_SRC=32768 timeout 600 python tools/gpu_time.py C5 global bfs_words=2 pull_alpha=$A
  done
  echo "== C5 slice words=2 pull_beta=2"; VGA_TIME_SRC=32768 timeout 600 python tools/gpu_time.py C5 global bfs_words=2 pull_beta=2
  for W in 1 2 4 8; do
    echo "== C2 words=$W"; timeout 300 python tools/gpu_time.py C2 global bfs_words=$W
  done
  echo "== C4 slice words=4"; VGA_TIME_SRC=16384 timeout 600 python tools/gpu_time.py C4 global bfs_words=4
  echo "== C4 slice words=8"; VGA_TIME_SRC=16384 timeout 600 python tools/gpu_time.py C4 global bfs_words=8
  echo "== C1"; timeout 300 python tools/gpu_time.py C1 global
} > gpurun_out/r2c3_ab.log 2>&1
export VGA_TIME_SRC=8192
CMD="python tools/gpu_time.py C5 global bfs_words=2"
$CMD > gpurun_out/r2c3_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c3_launches_C5slice.csv $CMD > gpurun_out/r2c3_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in k_push_nodes k_pull_nodes k_update; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 2 -o gpurun_out/r2c3_prof_$K $CMD > gpurun_out/r2c3_ncu_$K.log 2>&1
  echo "$K capture rc=$?"
  python tools/ncu_summary.py gpurun_out/r2c3_prof_$K.ncu-rep > gpurun_out/r2c3_prof_${K}_summary.txt 2>/dev/null
done
ls -la gpurun_out | tail -30
