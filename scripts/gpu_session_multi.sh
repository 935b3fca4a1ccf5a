#!/bin/bash
set -x
N=${1:-2}
S=${2:-m2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_fused_allreduce_gpu.py -m gpu -q -s > gpurun_out/${S}_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/${S}_pytest_multi.log
tail -8 gpurun_out/${S}_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${S}_bench_${N}gpu.json 2> gpurun_out/${S}_bench_${N}gpu.err
tail -c 2500 gpurun_out/${S}_bench_${N}gpu.json; tail -8 gpurun_out/${S}_bench_${N}gpu.err
