#!/bin/bash
# other solvers (DP5 / BS3 / RK4): parity tests, Tsit5 regression (tests + bench)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_solvers_gpu.py -q -m gpu -x -s > gpurun_out/s11_solvers.log 2>&1; tail -25 gpurun_out/s11_solvers.log
timeout 1200 python -m pytest tests/test_goku_gpu.py tests/test_user_rhs_gpu.py tests/test_golden.py tests/test_abi.py -q -m gpu > gpurun_out/s11_goku.log 2>&1; tail -5 gpurun_out/s11_goku.log
timeout 600 python bench.py --no-cpu --no-training > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s11_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d.get("roofline"), d["discrete_adjoint"]["ms_per_step"], d["e2e"])
PY
