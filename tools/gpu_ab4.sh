#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_xbatch.py -m gpu -x -q -k "jpeg_share or distinct" > $O/ab4_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/ab4_tests.log
for ns in 0 1; do
  LP_HUFF_V=1 LP_HUFF_NOSTORE=$ns timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ab4_bench_ns$ns.json 2> $O/ab4_bench_ns$ns.err; echo "bench nostore=$ns rc=$?"
  python -c "
import json;d=json.load(open('$O/ab4_bench_ns$ns.json'));print('ns$ns',d['value'],d['config']['stage_ms_per_step'],d['config']['huffman_phase_share'])"
done
timeout 600 python bench.py --variant pcg64 --batch 1024 --steps 3 --warmup 3 --cpu-seconds 6 > $O/ab4_bench_pcg64.json 2> $O/ab4_bench_pcg64.err; echo "pcg64 rc=$?"; cut -c1-600 $O/ab4_bench_pcg64.json
