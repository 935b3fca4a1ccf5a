#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_properties_gpu.py -q -m gpu > gpurun_out/s23_props.log 2>&1; tail -4 gpurun_out/s23_props.log
timeout 1200 python -m pytest tests/test_goku_gpu.py tests/test_solvers_gpu.py tests/test_golden.py tests/test_user_rhs_gpu.py -q -m gpu > gpurun_out/s23_goku.log 2>&1; tail -3 gpurun_out/s23_goku.log
timeout 600 python bench.py --no-cpu --no-training --steps 20 > gpurun_out/s23_bench.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s23_bench.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "DA", d["discrete_adjoint"]["ms_per_step"], d["discrete_adjoint"]["fwd_ms"], d["discrete_adjoint"]["bwd_ms"])
PY
