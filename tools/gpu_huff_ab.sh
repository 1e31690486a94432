#!/bin/bash
# A/B of the JPEG entropy-decode kernel variants (LP_HUFF_V) + the parity tests that drive them.
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_random_differential.py tests/test_jpeg_optimized_tables.py tests/test_jpeg_multiscan.py tests/test_resize_cubic.py tests/test_gpu_xbatch.py -m gpu -x -q > $O/ab_tests_v2.log 2>&1; echo "tests v2 rc=$?"; tail -3 $O/ab_tests_v2.log
for v in 3 4 5; do
  LP_HUFF_V=$v timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_jpeg_optimized_tables.py tests/test_gpu_random_differential.py -m gpu -x -q > $O/ab_tests_v$v.log 2>&1; echo "tests v$v rc=$?"; tail -1 $O/ab_tests_v$v.log
done
for v in 1 2 3 4 5; do
  LP_HUFF_V=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ab_bench_v$v.json 2> $O/ab_bench_v$v.err; echo "bench v$v rc=$?"
  python -c "
import json;d=json.load(open('$O/ab_bench_v$v.json'));print('v$v',d['value'],d['ms_per_step'],d['config']['stage_ms_per_step'],d['e2e']['value'],d['config']['huffman_sync_rounds'])"
done
