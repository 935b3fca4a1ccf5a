#!/bin/bash
# round-2 re-entry: full GPU suite + default bench + reference arm on a fresh box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s10_smi.txt
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/s10_pytest.log 2>&1; tail -5 gpurun_out/s10_pytest.log
timeout 600 python bench.py > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; tail -c 1500 gpurun_out/s10_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s10_ref.json 2> gpurun_out/s10_ref.err; tail -c 600 gpurun_out/s10_ref.json
