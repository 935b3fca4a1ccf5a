#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "interpolating" > gpurun_out/s17_cadj_tests.log 2>&1; tail -4 gpurun_out/s17_cadj_tests.log
LDEQ_CADJ_BATCH=0 timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/s17_c2_rec0.json 2>/dev/null
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/s17_c2_rec1.json 2>/dev/null
python - <<'PY'
import json
for f in ("s17_c2_rec0","s17_c2_rec1"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    v=d["variants"]["exact_fp32_global_interpolating_adjoint"]
    print(f, {k:(round(x,3) if isinstance(x,float) else x) for k,x in v.items() if k in ("ms","fwd_bwd_ms","bwd_ms","bwd_steps","naccept_mean","bwd_stats")})
PY
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "interpolating_adjoint_fp32_c2" > gpurun_out/s17_race.log 2>&1; tail -2 gpurun_out/s17_race.log
