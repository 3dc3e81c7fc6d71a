#!/bin/bash
# hqp contact-columns-once + warm stage A: parity at full size, bench, phases
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r02b_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/r02b_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > $O/r02b_bench_c2.json 2> $O/r02b_bench_c2.err
timeout 400 python bench.py --config 4 --steps 10 --no-cpu-baseline > $O/r02b_bench_c4.json 2> $O/r02b_bench_c4.err
timeout 400 python bench.py --config 3i --steps 10 --no-cpu-baseline > $O/r02b_bench_c3i.json 2> $O/r02b_bench_c3i.err
timeout 400 python bench.py --outputs all --steps 10 --no-cpu-baseline > $O/r02b_bench_c2_all.json 2> $O/r02b_bench_c2_all.err
tail -5 $O/r02b_pytest_gpu.log
python tools/bench_brief.py $O/r02b_bench_c2.json $O/r02b_bench_c4.json $O/r02b_bench_c3i.json $O/r02b_bench_c2_all.json
