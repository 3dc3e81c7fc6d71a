#!/bin/bash
# A/B of a library variant on config 4 (mixed support) with the parity tests: bash tools/gpu_ab_c4.sh name
mkdir -p gpurun_out
O=gpurun_out
v=$1
DWBC_LIB=$PWD/tools/_build/$v.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or lean or config2 or golden or walking or perturbed" > $O/ab_parity_$v.log 2>&1
tail -2 $O/ab_parity_$v.log
for rep in 1 2; do
  timeout 400 python bench.py --config 4 --steps 10 --no-cpu-baseline > $O/abc4_base_$rep.json 2>/dev/null
  DWBC_LIB=$PWD/tools/_build/$v.so timeout 400 python bench.py --config 4 --steps 10 --no-cpu-baseline > $O/abc4_${v}_$rep.json 2>/dev/null
done
DWBC_LIB=$PWD/tools/_build/$v.so timeout 400 python bench.py --steps 10 --no-cpu-baseline --contact-mix ss > $O/abc4_${v}_ss.json 2>/dev/null
timeout 400 python bench.py --steps 10 --no-cpu-baseline --contact-mix ss > $O/abc4_base_ss.json 2>/dev/null
python tools/bench_brief.py $O/abc4_*.json
