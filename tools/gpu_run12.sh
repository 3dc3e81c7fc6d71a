#!/bin/bash
set -x
mkdir -p gpurun_out
O=gpurun_out
for v in base gang2 notz; do
  if [ $v = base ]; then unset DWBC_LIB; else export DWBC_LIB=$PWD/tools/_build/$v.so; fi
  timeout 400 python bench.py --steps 10 --no-cpu-baseline > $O/r02f_bench_$v.json 2> $O/r02f_bench_$v.err
done
unset DWBC_LIB
python tools/bench_brief.py $O/r02f_bench_base.json $O/r02f_bench_gang2.json $O/r02f_bench_notz.json
