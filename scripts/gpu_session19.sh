#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrent_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/s19_tests.log 2>&1; tail -6 gpurun_out/s19_tests.log
for tool in memcheck racecheck; do timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py recurrent > gpurun_out/s19_${tool}.log 2>&1; grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/s19_${tool}.log | tail -1; done
timeout 300 python scripts/quick_recurrent.py 8192 2>/dev/null
timeout 600 python examples/pendulum_train.py --model latentode --epochs 2 2>&1 | tail -3
