#!/bin/bash
# GPU session 1 of round 2: parity tests, bench in both output groups, launch list, full ncu capture of the two hot kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_torques.json 2> gpurun_out/bench_torques.err
timeout 300 python bench.py --steps 10 --warmup 3 --outputs all --no-cpu-baseline > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err
timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix mixed --batch 131072 --no-cpu-baseline > gpurun_out/bench_mixed.json 2> gpurun_out/bench_mixed.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lean_cycle_kernel|hqp_kernel' -s 2 -c 2 -f -o gpurun_out/prof_r2a python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2a.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_torques.json | head -c 1500
