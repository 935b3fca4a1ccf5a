#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_properties_gpu.py -q -m gpu > gpurun_out/s22_props.log 2>&1; tail -15 gpurun_out/s22_props.log
