#!/bin/bash
set -x
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02e}
timeout 1500 python -m pytest tests -m gpu -q -x > $O/${T}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/${T}_pytest_gpu.log
timeout 400 python bench.py --no-cpu-baseline > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
timeout 400 python bench.py --config 4 --steps 10 --no-cpu-baseline > $O/${T}_bench_c4.json 2> $O/${T}_bench_c4.err
timeout 400 python bench.py --config 0 --steps 10 --no-cpu-baseline > $O/${T}_bench_c0.json 2> $O/${T}_bench_c0.err
timeout 300 python tools/lean_phases.py 65536 > $O/${T}_phases.json 2> $O/${T}_phases.err
tail -4 $O/${T}_pytest_gpu.log
python tools/bench_brief.py $O/${T}_bench_c2.json $O/${T}_bench_c4.json $O/${T}_bench_c0.json
head -28 $O/${T}_phases.json
