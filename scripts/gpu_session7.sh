#!/bin/bash
# Round-2 evidence run: full GPU suite, smoke, ncu captures of every hot kernel, launch list of the bench command, bench line.
set -x
S=${1:-s7}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${S}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${S}_pytest.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/${S}_pytest.log | tail -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${S}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${S}_smoke.log
timeout 300 python scripts/prof_loss.py > gpurun_out/${S}_loss.json 2> gpurun_out/${S}_loss.err; cat gpurun_out/${S}_loss.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'elbo' -c 6 -f -o gpurun_out/r2_${S}_loss python scripts/prof_loss.py > gpurun_out/${S}_ncu_loss.log 2>&1; tail -2 gpurun_out/${S}_ncu_loss.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tsit5_(fwd|bwd|fwdsens)' -s 6 -c 5 -f -o gpurun_out/r2_${S}_goku python scripts/prof_goku.py > gpurun_out/${S}_ncu.log 2>&1; tail -2 gpurun_out/${S}_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_tc_(adj|wgrad|fwd)' -s 6 -c 4 -f -o gpurun_out/r2_${S}_tc python scripts/prof_mlp_tc.py > gpurun_out/${S}_ncu_tc.log 2>&1; tail -2 gpurun_out/${S}_ncu_tc.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_${S}_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-training > gpurun_out/${S}_bench_under_ncu.log 2>&1; wc -l gpurun_out/r2_${S}_bench_launches.csv
timeout 900 python bench.py > gpurun_out/${S}_bench.json 2> gpurun_out/${S}_bench.err; cat gpurun_out/${S}_bench.json; tail -2 gpurun_out/${S}_bench.err
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/${S}_bench_c2.json 2> gpurun_out/${S}_bench_c2.err; cat gpurun_out/${S}_bench_c2.json; tail -2 gpurun_out/${S}_bench_c2.err
