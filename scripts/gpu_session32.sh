#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tsit5_(fwd|bwd|fwdsens)' -s 6 -c 5 -f -o gpurun_out/r2_final_goku python scripts/prof_goku.py > gpurun_out/s32_ncu.log 2>&1; tail -2 gpurun_out/s32_ncu.log
