#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu -s -k "interpolating" > gpurun_out/s9_pytest.log 2>&1; echo "rc=$?"
grep -E "backward solve|passed|failed|Error|error|assert|^E " gpurun_out/s9_pytest.log | tail -30
