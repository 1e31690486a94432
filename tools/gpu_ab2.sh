#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_webp_encode.py tests/test_gpu_xbatch.py tests/test_gpu_webp.py -m gpu -x -q > $O/ab2_tests_webp.log 2>&1; echo "webp tests rc=$?"; tail -3 $O/ab2_tests_webp.log
for v in 1 3; do
  LP_HUFF_V=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ab2_bench_v$v.json 2> $O/ab2_bench_v$v.err; echo "bench v$v rc=$?"
  python -c "
import json;d=json.load(open('$O/ab2_bench_v$v.json'));print('v$v',d['value'],d['config']['stage_ms_per_step'],d['config']['huffman_phase_share'])"
done
