#!/bin/bash
# the lossy WebP encoder after a change: device stream == serial core (tests), configs 4 and 3 timed
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_webp_encode.py tests/test_gpu_xbatch.py tests/test_gpu_webp.py -m gpu -x -q > $O/enc_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/enc_tests.log
for c in 4 3; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > $O/enc_bench_c$c.json 2> $O/enc_bench_c$c.err; echo "bench c$c rc=$?"
  python -c "
import json;d=json.load(open('$O/enc_bench_c$c.json'));print('c$c',d['value'],d['e2e']['value'],d['config']['stage_ms_per_step'])"
done
