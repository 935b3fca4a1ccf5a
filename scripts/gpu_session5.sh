#!/bin/bash
set -x
S=${1:-s5}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py tests/test_golden.py tests/test_abi.py -m gpu -q -s -x > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${S}_pytest.log
grep -E "passed|failed|FAILED|tc bwd|Error|error" gpurun_out/${S}_pytest.log | tail -20
timeout 600 python scripts/quick_mlp_tc.py > gpurun_out/${S}_mlp_tc.json 2> gpurun_out/${S}_mlp_tc.err; cat gpurun_out/${S}_mlp_tc.json; tail -3 gpurun_out/${S}_mlp_tc.err
