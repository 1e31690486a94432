#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_png.py tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_jpeg_optimized_tables.py -m gpu -x -q > $O/ab3_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/ab3_tests.log
for v in 6 7 8; do
  LP_HUFF_V=$v timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_jpeg_optimized_tables.py -m gpu -x -q > $O/ab3_tests_v$v.log 2>&1; echo "tests v$v rc=$?"
done
for v in 1 2 6 7 8; do
  LP_HUFF_V=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ab3_bench_v$v.json 2> $O/ab3_bench_v$v.err; echo "bench v$v rc=$?"
  python -c "
import json;d=json.load(open('$O/ab3_bench_v$v.json'));print('v$v',d['value'],d['config']['stage_ms_per_step']['huff_decode'],d['config']['huffman_phase_share'])"
done
