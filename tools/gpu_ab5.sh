#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_webp_encode.py tests/test_gpu_xbatch.py tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -x -q > $O/ab5_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/ab5_tests.log
for c in 4 3; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 2 --no-cpu-baseline > $O/ab5_bench_c$c.json 2> $O/ab5_bench_c$c.err; echo "bench c$c rc=$?"
  python -c "
import json;d=json.load(open('$O/ab5_bench_c$c.json'));print('c$c',d['value'],d['e2e']['value'],d['config']['stage_ms_per_step'])"
done
