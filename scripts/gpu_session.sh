#!/bin/bash
# One GPU session: tests, timing (A/B variants), ncu captures.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/s1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s1_pytest.log
tail -30 gpurun_out/s1_pytest.log
timeout 300 python scripts/quick_goku.py > gpurun_out/s1_quick.json 2> gpurun_out/s1_quick.err; cat gpurun_out/s1_quick.json
LDEQ_LIB=$PWD/latentdiffeq.jl_b200/lib/variants/libldeq_nof32x2.so timeout 300 python scripts/quick_goku.py > gpurun_out/s1_quick_nof32x2.json 2>> gpurun_out/s1_quick.err; cat gpurun_out/s1_quick_nof32x2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tsit5_(fwd|bwd|fwdsens)' -s 6 -c 5 -f -o gpurun_out/r2_s1_goku python scripts/prof_goku.py > gpurun_out/s1_ncu.log 2>&1
tail -5 gpurun_out/s1_ncu.log
ls -la gpurun_out | tail -20
