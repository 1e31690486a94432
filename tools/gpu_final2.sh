#!/bin/bash
O=gpurun_out; mkdir -p $O
for c in 2 3 4 5; do
  timeout 700 python bench.py --config $c --steps 3 --warmup 3 > $O/r02g_bench_c$c.json 2> $O/r02g_bench_c$c.err; echo "bench c$c rc=$?"
  python -c "
import json;d=json.load(open('$O/r02g_bench_c$c.json'));print('c$c',d['value'],d['e2e']['value'],d['cpu_baseline']['value'],d['cpu_baseline']['sample'][:90])"
done
