#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py solvers > gpurun_out/s16_${tool}_solvers.log 2>&1; echo "$tool solvers: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/s16_${tool}_solvers.log | tail -1)"
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py recurrent > gpurun_out/s16_${tool}_recurrent.log 2>&1; echo "$tool recurrent: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/s16_${tool}_recurrent.log | tail -1)"
done
timeout 600 python examples/pendulum_train.py --model goku --epochs 2 > gpurun_out/s16_train_goku.log 2>&1; tail -4 gpurun_out/s16_train_goku.log
timeout 600 python examples/pendulum_train.py --model latentode --epochs 2 > gpurun_out/s16_train_latentode.log 2>&1; tail -4 gpurun_out/s16_train_latentode.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 ./tests/cabi_smoke | tail -2
