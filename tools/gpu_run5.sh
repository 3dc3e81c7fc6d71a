#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/status_mismatch_*.json gpurun_out/active_mismatch_*.json
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "full_size_every_state or lean_equals or reduced_lqp" > gpurun_out/pytest_full.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_full.log
for mix in ds ss sorted mixed; do
  timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix $mix --no-cpu-baseline > gpurun_out/bench_mix_$mix.json 2> gpurun_out/bench_mix_$mix.err
done
DWBC_LIB=$PWD/scratch/libdwbc_rollk.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rollk.json 2> gpurun_out/bench_rollk.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lean_cycle_kernel' -s 2 -c 1 -f -o gpurun_out/prof_r2c_mixed python bench.py --steps 1 --warmup 1 --contact-mix mixed --no-cpu-baseline > gpurun_out/ncu_r2c.log 2>&1
tail -5 gpurun_out/pytest_full.log
