#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err; tail -c 600 gpurun_out/s14_bench.json; tail -3 gpurun_out/s14_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe_ -s 6 -c 6 -o gpurun_out/s14_pe python scripts/prof_recurrent.py > gpurun_out/s14_ncu.log 2>&1; tail -3 gpurun_out/s14_ncu.log
