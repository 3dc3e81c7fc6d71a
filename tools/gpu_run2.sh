#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/lean_phases.py 65536 > gpurun_out/phases_ds.json 2> gpurun_out/phases_ds.err
timeout 300 python tools/lean_phases.py 65536 --mixed > gpurun_out/phases_mixed.json 2> gpurun_out/phases_mixed.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_torques.json 2> gpurun_out/bench_torques.err
timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix mixed --batch 131072 --no-cpu-baseline > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
tail -5 gpurun_out/pytest_gpu.log
