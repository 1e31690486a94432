#!/bin/bash
# See host_threads_fake_gpu.cpp: the library under ThreadSanitizer over the pretend CUDA runtime, many threads.
# CPU only; seeds from tests/native/host_parse_fuzz.sh (run it once first).
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
HERE=$ROOT/lilliput_b200/csrc
OUT=${LP_TSAN_DIR:-/tmp/tsan}
SEEDS=${LP_ASAN_DIR:-/tmp/asan}/seeds
[ -n "$(ls $SEEDS 2>/dev/null)" ] || { echo "run tests/native/host_parse_fuzz.sh once first (it dumps the seed files)"; exit 2; }
mkdir -p $OUT/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=thread,-fno-omit-frame-pointer -I$ROOT/include -I$ROOT/lilliput_b200/host -I$HERE"
cd $OUT/build
for f in resize jpeg_decode jpeg_huff_parallel jpeg_encode png_decode png_encode gif_decode webp_decode webp_encode pixel_ops abi_opencv batch; do
  [ $f.o -nt $HERE/$f.cu ] || ( nvcc $FLAGS -c $HERE/$f.cu -o $f.o ) &
done
[ jpeg_parse.o -nt $HERE/jpeg_parse.cpp ] || ( nvcc $FLAGS -x cu -c $HERE/jpeg_parse.cpp -o jpeg_parse.o ) &
[ png_parse.o -nt $HERE/png_parse.cpp ] || ( nvcc $FLAGS -x cu -c $HERE/png_parse.cpp -o png_parse.o ) &
[ lilliput_host.o -nt $ROOT/lilliput_b200/host/lilliput_host.cpp ] || ( nvcc $FLAGS -x cu -c $ROOT/lilliput_b200/host/lilliput_host.cpp -o lilliput_host.o ) &
wait
g++ -O1 -g -std=c++17 -fPIC -fsanitize=thread -I${CUDA_HOME:-/usr/local/cuda}/include -c $ROOT/tests/native/fake_cudart.cpp -o fake_cudart.o
g++ -shared -fsanitize=thread -o $OUT/liblp_tsan.so *.o
cd $OUT
g++ -O1 -g -std=c++17 -fsanitize=thread -I$ROOT/include $ROOT/tests/native/host_threads_fake_gpu.cpp -o threads_fake -L. -llp_tsan -Wl,-rpath,$OUT -lpthread
./threads_fake ${1:-16} ${2:-5000} $SEEDS/* 2>&1 | grep -v "^\[lilliput" | tail -40
