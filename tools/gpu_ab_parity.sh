#!/bin/bash
# A/B with the full-size parity tests on the variant: bash tools/gpu_ab_parity.sh name
mkdir -p gpurun_out
O=gpurun_out
v=$1
DWBC_LIB=$PWD/tools/_build/$v.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or lean or config2 or golden or walking or perturbed" > $O/ab_parity_$v.log 2>&1
tail -3 $O/ab_parity_$v.log
bash tools/gpu_ab.sh $v
