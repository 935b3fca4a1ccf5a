#!/bin/bash
mkdir -p gpurun_out
LDEQ_FWDSENS_SORT=0 python scripts/quick_goku.py > gpurun_out/s6_nosort.json 2> gpurun_out/s6_nosort.err
python scripts/quick_goku.py > gpurun_out/s6_sort.json 2> gpurun_out/s6_sort.err
LDEQ_LIB=$PWD/latentdiffeq.jl_b200/lib/libldeq_t256.so python scripts/quick_goku.py > gpurun_out/s6_sort256.json 2> gpurun_out/s6_sort256.err
timeout 900 python -m pytest tests/test_goku_gpu.py -x -q -m gpu > gpurun_out/s6_pytest.log 2>&1
tail -3 gpurun_out/s6_pytest.log
python - <<'PY'
import json
for n in ("nosort","sort","sort256"):
    try:
        d=json.load(open(f"gpurun_out/s6_{n}.json")); print(n, d["fwddual"]["bwd_ms"], d["fwddual"]["fwd_bwd_ms"])
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/s6_{n}.err").read()[-500:])
PY
