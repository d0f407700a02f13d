#!/bin/bash
# GPU call 15 of round 2 (1 GPU): metric / angular VGA (row f4): parity tests, CLI drop-in cases, then timing on C1 / C2.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_metric.py tests/test_cli_dropin.py -m gpu -x -q -p no:cacheprovider -k "metric or angular or cli_outputs" > gpurun_out/r2c15_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2c15_pytest.log
timeout 600 python tools/metric_time.py C1 2048 16 > gpurun_out/r2c15_metric_C1.log 2>&1; echo "C1 rc=$?"; cat gpurun_out/r2c15_metric_C1.log
timeout 900 python tools/metric_time.py C2 4096 8 > gpurun_out/r2c15_metric_C2.log 2>&1; echo "C2 rc=$?"; cat gpurun_out/r2c15_metric_C2.log
