#!/bin/bash
# last check of the final tree: full GPU suite + smoke + C smoke + a short default bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/s31_pytest.log 2>&1; tail -3 gpurun_out/s31_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 ./tests/cabi_smoke | tail -1
timeout 900 python bench.py --steps 20 > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s31_bench.json").read().strip().splitlines()[-1])
print("value",d["value"],d["ms_per_step"],"e2e",d["e2e"]["value"],d["e2e"]["link_roofline"]["frac"],"training",d["training"]["value"],"cadj",d["latentode"]["c2_batch_256"]["forward_interpolating_adjoint_ms"])
PY
