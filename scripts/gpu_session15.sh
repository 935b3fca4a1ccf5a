#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrent_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/s15_tests.log 2>&1; tail -4 gpurun_out/s15_tests.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_recurrent_gpu.py::test_forward_and_gradients_match_the_oracle[37-7-32]" -q -m gpu > gpurun_out/s15_racecheck.log 2>&1; tail -2 gpurun_out/s15_racecheck.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_recurrent_gpu.py::test_forward_and_gradients_match_the_oracle[37-7-32]" "tests/test_recurrent_gpu.py::test_latentode_rnn_only" -q -m gpu > gpurun_out/s15_memcheck.log 2>&1; tail -2 gpurun_out/s15_memcheck.log
timeout 300 python scripts/quick_recurrent.py 8192 > gpurun_out/s15_recurrent_timing.json 2>/dev/null; cat gpurun_out/s15_recurrent_timing.json
timeout 300 python scripts/quick_recurrent.py 65536 > gpurun_out/s15_recurrent_timing_64k.json 2>/dev/null; cat gpurun_out/s15_recurrent_timing_64k.json
timeout 900 python bench.py --workload c5 --global-batch 8192 --steps 8 --warmup 3 --no-cpu > gpurun_out/s15_c5_8192.json 2> gpurun_out/s15_c5.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s15_c5_8192.json").read().strip().splitlines()[-1])
print(d["training"]["nccl"])
PY
