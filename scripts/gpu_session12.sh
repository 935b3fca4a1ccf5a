#!/bin/bash
mkdir -p gpurun_out
( lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; numactl -H 2>/dev/null | head -20; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo $d $(cat $d/numa_node) $(cat $d/class); fi; done; free -g | head -2 ) > gpurun_out/s12_topology.txt 2>&1
timeout 900 python -m pytest tests/test_recurrent_gpu.py -q -m gpu -x > gpurun_out/s12_recurrent.log 2>&1; tail -15 gpurun_out/s12_recurrent.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest "tests/test_recurrent_gpu.py::test_forward_and_gradients_match_the_oracle[37-7-32]" -q -m gpu > gpurun_out/s12_memcheck.log 2>&1; tail -4 gpurun_out/s12_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest "tests/test_recurrent_gpu.py::test_forward_and_gradients_match_the_oracle[37-7-32]" -q -m gpu > gpurun_out/s12_racecheck.log 2>&1; tail -4 gpurun_out/s12_racecheck.log
timeout 300 python scripts/quick_recurrent.py 8192 > gpurun_out/s12_recurrent_timing.json 2> gpurun_out/s12_recurrent_timing.err; cat gpurun_out/s12_recurrent_timing.json; tail -2 gpurun_out/s12_recurrent_timing.err
timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu > gpurun_out/s12_model.log 2>&1; tail -3 gpurun_out/s12_model.log
