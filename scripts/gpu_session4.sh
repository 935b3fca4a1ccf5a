#!/bin/bash
set -x
S=${1:-s4}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${S}_pytest.log
grep -E "passed|failed|FAILED" gpurun_out/${S}_pytest.log | tail -12
timeout 300 python scripts/quick_goku.py > gpurun_out/${S}_quick.json 2> gpurun_out/${S}_quick.err; cat gpurun_out/${S}_quick.json
timeout 300 python scripts/prof_loss.py > gpurun_out/${S}_loss.json 2> gpurun_out/${S}_loss.err; cat gpurun_out/${S}_loss.json; tail -3 gpurun_out/${S}_loss.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'elbo|adamw|sample' -c 12 -f -o gpurun_out/r2_${S}_loss python scripts/prof_loss.py > gpurun_out/${S}_ncu_loss.log 2>&1; tail -3 gpurun_out/${S}_ncu_loss.log
# launch list of one C5 training step (all kernels, incl. torch / cuDNN / cuBLAS ones): warm-up launches skipped by count
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "c5_step/" --csv --log-file gpurun_out/r2_${S}_c5_launches.csv python scripts/prof_c5_step.py > gpurun_out/${S}_ncu_c5.log 2>&1; tail -3 gpurun_out/${S}_ncu_c5.log; wc -l gpurun_out/r2_${S}_c5_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tsit5_(fwd|bwd|fwdsens)' -s 6 -c 5 -f -o gpurun_out/r2_${S}_goku python scripts/prof_goku.py > gpurun_out/${S}_ncu.log 2>&1
# launch list of the bench command itself (short)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_${S}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-training > gpurun_out/${S}_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_${S}_bench_launches.csv
