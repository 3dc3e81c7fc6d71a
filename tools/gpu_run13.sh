#!/bin/bash
# final round-2 validation: smoke, full GPU suite, every configuration, the reference arm, ncu launch list and full capture of one device-resident step
set -x
mkdir -p gpurun_out
O=gpurun_out
T=${TAG:-r02final}
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1
timeout 2400 python -m pytest tests -m gpu -q > $O/${T}_pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/${T}_pytest_gpu.log
timeout 400 python bench.py > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err
for c in 0 1 3i 3ii 4; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > $O/${T}_bench_c$c.json 2> $O/${T}_bench_c$c.err
done
timeout 400 python bench.py --outputs all --steps 10 > $O/${T}_bench_c2_all.json 2> $O/${T}_bench_c2_all.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref_c2.json 2> $O/${T}_bench_ref_c2.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${T}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lean_cycle_kernel|hqp_kernel|lean_classify' -s 5 -c 5 -f -o $O/${T}_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/${T}_ncu_full.log 2>&1
DWBC_LIB=$PWD/tools/_build/timing.so timeout 300 python tools/lean_phases.py 65536 > $O/${T}_phases.json 2> $O/${T}_phases.err
tail -4 $O/${T}_pytest_gpu.log; tail -2 $O/${T}_smoke.log
python tools/bench_brief.py $O/${T}_bench_c2.json $O/${T}_bench_c0.json $O/${T}_bench_c1.json $O/${T}_bench_c3i.json $O/${T}_bench_c3ii.json $O/${T}_bench_c4.json $O/${T}_bench_c2_all.json
