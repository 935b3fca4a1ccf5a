#!/bin/bash
set -x
S=${1:-s3}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${S}_pytest.log
grep -E "passed|failed|FAILED" gpurun_out/${S}_pytest.log | tail -12
timeout 300 python scripts/quick_goku.py > gpurun_out/${S}_quick.json 2> gpurun_out/${S}_quick.err; cat gpurun_out/${S}_quick.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; tail -c 3000 gpurun_out/${S}_bench.json; tail -5 gpurun_out/${S}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${S}_bench_ref.json 2> gpurun_out/${S}_bench_ref.err; tail -c 1500 gpurun_out/${S}_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${S}_smoke.log 2>&1; tail -3 gpurun_out/${S}_smoke.log
