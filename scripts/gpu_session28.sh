#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do export LDEQ_FWD_SORT=$v;
  timeout 600 python bench.py --no-cpu --no-training --steps 20 > gpurun_out/s28_bench_$v.json 2>/dev/null
  python - $v <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/s28_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
da=d["discrete_adjoint"]
print("fwd sort",sys.argv[1],"headline ms",round(d["ms_per_step"],4),"DA ms",round(da["ms_per_step"],4),"fwd",round(da["fwd_ms"],4),"bwd",round(da["bwd_ms"],4))
PY
done
LDEQ_FWD_SORT=1 timeout 1200 python -m pytest tests/test_properties_gpu.py tests/test_goku_gpu.py tests/test_solvers_gpu.py tests/test_user_rhs_gpu.py tests/test_golden.py -q -m gpu > gpurun_out/s28_tests.log 2>&1; tail -3 gpurun_out/s28_tests.log
LDEQ_FWD_SORT=1 timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize_small.py goku > gpurun_out/s28_race.log 2>&1; tail -1 gpurun_out/s28_race.log
LDEQ_FWD_SORT=1 timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py goku > gpurun_out/s28_mem.log 2>&1; tail -1 gpurun_out/s28_mem.log
