#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/status_mismatch_*.json gpurun_out/active_mismatch_*.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for mix in ds ss sorted mixed; do
  timeout 300 python bench.py --steps 10 --warmup 3 --contact-mix $mix --no-cpu-baseline > gpurun_out/bench_mix_$mix.json 2> gpurun_out/bench_mix_$mix.err
done
timeout 300 python tools/lean_phases.py 65536 > gpurun_out/phases_ds.json 2> gpurun_out/phases_ds.err
tail -5 gpurun_out/pytest_gpu.log
