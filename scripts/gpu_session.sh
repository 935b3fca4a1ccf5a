#!/bin/bash
# One GPU session: tests, timing (A/B variants), ncu captures.  Outputs under gpurun_out/.
set -x
S=${1:-s2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${S}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${S}_pytest.log
tail -40 gpurun_out/${S}_pytest.log
timeout 300 python scripts/quick_goku.py > gpurun_out/${S}_quick.json 2> gpurun_out/${S}_quick.err; cat gpurun_out/${S}_quick.json
for v in nof32x2 occ65 nopf; do
  LDEQ_LIB=$PWD/latentdiffeq.jl_b200/lib/variants/libldeq_$v.so timeout 300 python scripts/quick_goku.py > gpurun_out/${S}_quick_$v.json 2>> gpurun_out/${S}_quick.err; cat gpurun_out/${S}_quick_$v.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tsit5_(fwd|bwd|fwdsens)' -s 6 -c 5 -f -o gpurun_out/r2_${S}_goku python scripts/prof_goku.py > gpurun_out/${S}_ncu.log 2>&1
tail -5 gpurun_out/${S}_ncu.log
ls -la gpurun_out | tail -20
