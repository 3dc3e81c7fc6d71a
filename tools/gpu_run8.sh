#!/bin/bash
# round-2 validation session: smoke, full GPU suite, every bench configuration, reference arm, ncu launch list + full capture
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r02_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/r02_pytest_gpu.log
timeout 400 python bench.py > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
for c in 0 1 3i 3ii 4; do
  timeout 400 python bench.py --config $c --steps 10 --warmup 3 > $O/r02_bench_c$c.json 2> $O/r02_bench_c$c.err
done
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref_c2.json 2> $O/r02_bench_ref_c2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lean_cycle_kernel|hqp_kernel|lean_classify' -s 8 -c 4 -f -o $O/r02_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/r02_ncu_full.log 2>&1
tail -5 $O/r02_pytest_gpu.log; cat $O/r02_smoke.log | tail -3
