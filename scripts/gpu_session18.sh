#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "interpolating" -s > gpurun_out/s18_cadj_tests.log 2>&1; grep "backward solve" gpurun_out/s18_cadj_tests.log; tail -3 gpurun_out/s18_cadj_tests.log
LDEQ_CADJ_RES=0 timeout 900 python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "interpolating" -s > gpurun_out/s18_cadj_tests_nores.log 2>&1; grep "backward solve" gpurun_out/s18_cadj_tests_nores.log; tail -2 gpurun_out/s18_cadj_tests_nores.log
for v in "LDEQ_CADJ_BATCH=0" "LDEQ_CADJ_RES=0" "X=1"; do
  env $v timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/s18_c2.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
d=json.load(open("gpurun_out/s18_c2.json"))
v=d["variants"]["exact_fp32_global_interpolating_adjoint"]
print(sys.argv[1], {k:(round(x,3) if isinstance(x,float) else x) for k,x in v.items() if k in ("ms","fwd_bwd_ms")})
PY
done
cp gpurun_out/s18_c2.json gpurun_out/s18_c2_final.json
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py mlp > gpurun_out/s18_race_mlp.log 2>&1; tail -2 gpurun_out/s18_race_mlp.log
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_small.py mlp > gpurun_out/s18_mem_mlp.log 2>&1; tail -2 gpurun_out/s18_mem_mlp.log
