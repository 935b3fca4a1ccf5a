#!/bin/bash
# final 1-GPU validation: full GPU suite, smoke, default bench, reference arm, launch list of the bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x > gpurun_out/s21_pytest.log 2>&1; tail -3 gpurun_out/s21_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; tail -c 300 gpurun_out/s21_bench.json; echo
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s21_ref.json 2> gpurun_out/s21_ref.err; tail -c 300 gpurun_out/s21_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s21_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-training > gpurun_out/s21_ncu_bench.log 2>&1
python scripts/launch_list_summary.py gpurun_out/s21_bench_launches.csv | head -8
