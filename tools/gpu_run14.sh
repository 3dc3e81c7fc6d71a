#!/bin/bash
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/time_host_slices.py > $O/r02_host_slices.json 2> $O/r02_host_slices.err
cat $O/r02_host_slices.json; tail -3 $O/r02_host_slices.err
DWBC_LIB=$PWD/tools/_build/hqp6.so timeout 400 python bench.py --steps 10 --no-cpu-baseline > $O/r02f_bench_hqp6.json 2> $O/r02f_bench_hqp6.err
python tools/bench_brief.py $O/r02f_bench_hqp6.json
