#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "lean" > gpurun_out/pytest_lean.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_lean.log
for mix in ds mixed; do
  timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix $mix --no-cpu-baseline > gpurun_out/bench_mix_$mix.json 2> gpurun_out/bench_mix_$mix.err
done
timeout 300 python tools/lean_phases.py 65536 > gpurun_out/phases_ds.json 2> gpurun_out/phases_ds.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lean_cycle_kernel|hqp_kernel' -s 5 -c 3 -f -o gpurun_out/prof_r2d python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r2d.log 2>&1
tail -3 gpurun_out/pytest_lean.log
