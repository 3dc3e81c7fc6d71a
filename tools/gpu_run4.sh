#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/status_mismatch_*.json gpurun_out/active_mismatch_*.json
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "full_size_every_state or lean_equals" > gpurun_out/pytest_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_full.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_torques.json 2> gpurun_out/bench_torques.err
tail -5 gpurun_out/pytest_full.log
