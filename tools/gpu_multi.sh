#!/bin/bash
# multi-GPU session: N = $1.  Weak scaling (with the strong appendix), strong scaling, single-process host path, config 4.
set -x
N=$1
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_bench_n${N}.json 2> $O/r02_bench_n${N}.err
timeout 600 $TR bench.py --gpus $N --steps 40 --warmup 5 --scaling strong > $O/r02_bench_n${N}_strong.json 2> $O/r02_bench_n${N}_strong.err
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --config 4 > $O/r02_bench_n${N}_c4.json 2> $O/r02_bench_n${N}_c4.err
timeout 600 python bench.py --gpus $N --single-process --steps 20 --warmup 3 > $O/r02_bench_n${N}_single.json 2> $O/r02_bench_n${N}_single.err
timeout 600 python bench.py --gpus $N --single-process --scaling strong --steps 40 --warmup 5 > $O/r02_bench_n${N}_single_strong.json 2> $O/r02_bench_n${N}_single_strong.err
python tools/bench_brief.py $O/r02_bench_n${N}.json $O/r02_bench_n${N}_strong.json $O/r02_bench_n${N}_c4.json
python - <<P
import json
for f in ("$O/r02_bench_n${N}_single.json", "$O/r02_bench_n${N}_single_strong.json"):
    try:
        d = json.load(open(f)); print(f, round(d["ms_per_step"], 2), "ms", round(d["value"]), d["latency_ms"], d["details"]["shard_ms_last"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
d = json.load(open("$O/r02_bench_n${N}.json")); print("strong appendix", d.get("strong_scaling"))
P
