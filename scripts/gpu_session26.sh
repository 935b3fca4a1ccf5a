#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_cadj -c 1 -o gpurun_out/s26_cadj python scripts/quick_cadj.py > gpurun_out/s26_ncu.log 2>&1; tail -2 gpurun_out/s26_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pe_ -s 3 -c 3 -o gpurun_out/s26_pe python scripts/prof_recurrent.py > gpurun_out/s26_ncu_pe.log 2>&1; tail -2 gpurun_out/s26_ncu_pe.log
