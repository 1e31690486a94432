#!/bin/bash
# One gpurun call that records the state of the round: GPU parity suite, the bench lines of configs 2-5,
# the reference arm, the ncu launch list and `--set full` summaries of the JPEG kernels.
# Usage (from the repo root on the GPU box): bash tools/gpu_round.sh [tag]
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_gputests.log
for c in 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 > $O/${TAG}_bench_c$c.json 2> $O/${TAG}_bench_c$c.err; echo "bench c$c rc=$?"
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_c2_reference.json 2> $O/${TAG}_bench_ref.err; echo "ref rc=$?"
for v in pcg64 optimized dri; do   # secondary corpora of config 2 (labelled in the JSON), 1024 images per step
  timeout 400 python bench.py --variant $v --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c2_$v.json 2> $O/${TAG}_bench_c2_$v.err; echo "bench c2 $v rc=$?"
done
K='regex:jpeg_huff_sync|jpeg_idct|jpeg_upsample_color|resize_area|jpeg_unstuff|jpeg_fdct|jpeg_entropy|compact_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file $O/${TAG}_launches_c2.csv \
  python bench.py --steps 1 --warmup 1 --batch 1332 --no-cpu-baseline > $O/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
K2='regex:jpeg_huff_sync|jpeg_idct_kernel|jpeg_upsample_color|resize_area|jpeg_unstuff'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K2" --launch-skip 0 -c 5 -o $O/${TAG}_jpeg_kernels -f \
  python bench.py --steps 1 --warmup 1 --batch 1332 --no-cpu-baseline > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
cat $O/${TAG}_bench_c*.json | cut -c1-400
