#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/status_mismatch_*.json
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/lean_phases.py 65536 > gpurun_out/phases_ds.json 2> gpurun_out/phases_ds.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_torques.json 2> gpurun_out/bench_torques.err
timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix mixed --batch 131072 --no-cpu-baseline > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'hqp_kernel' -s 2 -c 1 -f -o gpurun_out/prof_r2b python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2b.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
