#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu > gpurun_out/s30_mlp_tests.log 2>&1; tail -3 gpurun_out/s30_mlp_tests.log
timeout 600 python bench.py --workload c2 --no-cpu --steps 5 > gpurun_out/s30_c2.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open("gpurun_out/s30_c2.json"))
for k,v in d["variants"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms","fwd_bwd_ms","unavailable")})
PY
