#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/s9_bench_c2.json 2> gpurun_out/s9_bench_c2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s9_bench_c2.json"))
for k,v in d["variants"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms","fwd_bwd_ms","unavailable","naccept_mean")})
PY
timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/s9_pytest.log 2>&1; tail -3 gpurun_out/s9_pytest.log
