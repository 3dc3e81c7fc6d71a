#!/bin/bash
# A/B of library variants: bash tools/gpu_ab.sh name1 name2 ...   (base = the in-tree library)
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
for v in base "$@"; do
  if [ $v = base ]; then unset DWBC_LIB; else export DWBC_LIB=$PWD/tools/_build/$v.so; fi
  timeout 400 python bench.py --steps 20 --no-cpu-baseline > $O/ab_${v}_$rep.json 2> $O/ab_${v}_$rep.err
done
done
unset DWBC_LIB
python tools/bench_brief.py $O/ab_*_1.json $O/ab_*_2.json
