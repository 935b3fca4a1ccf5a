#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_model_gpu.py tests/test_abi.py -q -m gpu > gpurun_out/s29_tests.log 2>&1; tail -5 gpurun_out/s29_tests.log
timeout 900 python bench.py --workload c5 --global-batch 8192 --steps 8 --warmup 3 --no-cpu > gpurun_out/s29_c5_8192.json 2> gpurun_out/s29_c5.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s29_c5_8192.json").read().strip().splitlines()[-1])
print(d["training"]["nccl"])
PY
timeout 600 python examples/pendulum_train.py --model goku --epochs 2 2>&1 | tail -3
