#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_recurrent_gpu.py tests/test_model_gpu.py -q -m gpu > gpurun_out/s13_tests.log 2>&1; tail -4 gpurun_out/s13_tests.log
timeout 900 python bench.py --workload c5 --global-batch 8192 --steps 8 --warmup 3 --no-cpu > gpurun_out/s13_c5_8192.json 2> gpurun_out/s13_c5.err; tail -c 900 gpurun_out/s13_c5_8192.json; tail -3 gpurun_out/s13_c5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s13_c5_launches.csv python scripts/prof_c5_step.py > gpurun_out/s13_prof.log 2>&1; tail -3 gpurun_out/s13_prof.log
python scripts/launch_list_summary.py gpurun_out/s13_c5_launches.csv | head -40
