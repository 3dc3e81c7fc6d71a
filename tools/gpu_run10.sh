#!/bin/bash
set -x
mkdir -p gpurun_out
O=gpurun_out
DWBC_BENCH_HARD_FRAC=0 timeout 400 python bench.py --steps 10 --no-cpu-baseline > $O/r02c_bench_nohard.json 2> $O/r02c_bench_nohard.err
DWBC_BENCH_HARD_FRAC=0.2 timeout 400 python bench.py --steps 10 --no-cpu-baseline > $O/r02c_bench_hard20.json 2> $O/r02c_bench_hard20.err
python tools/bench_brief.py $O/r02c_bench_nohard.json $O/r02c_bench_hard20.json
