#!/bin/bash
mkdir -p gpurun_out
for v in "" t64 pf48 pf12; do
  if [ -n "$v" ]; then export LDEQ_LIB=$PWD/latentdiffeq.jl_b200/lib/libldeq_$v.so; fi
  timeout 600 python bench.py --no-cpu --no-training --steps 20 > gpurun_out/s33_bench_$v.json 2>/dev/null
  python - "$v" <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/s33_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1] or "default", "ms_per_step", round(d["ms_per_step"],3), "pullback launch_ms", round(d["roofline"]["launch_ms"],3))
PY
done
