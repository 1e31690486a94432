#!/bin/bash
O=gpurun_out; mkdir -p $O
for v in 0 1; do
  LP_WEBP_I4=$v timeout 600 python bench.py --config 4 --steps 2 --warmup 2 --no-cpu-baseline > $O/ab6_bench_c4_i4$v.json 2> $O/ab6_bench_c4_i4$v.err; echo "bench c4 i4=$v rc=$?"
  python -c "
import json;d=json.load(open('$O/ab6_bench_c4_i4$v.json'));print('c4 i4=$v',d['value'],d['e2e']['value'],d['config']['stage_ms_per_step'])"
done
