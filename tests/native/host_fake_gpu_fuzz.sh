#!/bin/bash
# The host side of the WHOLE library (per-image ABI through lp_transform, and the batch ABI) under ASan + UBSan with
# tests/native/fake_cudart.cpp standing in for the CUDA runtime: device memory is host memory, kernels do nothing.
# CPU only.  Run tests/native/host_parse_fuzz.sh once first (it compiles the sanitized objects and dumps the seeds).
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${LP_ASAN_DIR:-/tmp/asan}
ITERS=${1:-20000}
[ -f $OUT/build/batch.o ] || { echo "run tests/native/host_parse_fuzz.sh first"; exit 2; }
mkdir -p $OUT/fake
cd $OUT/fake
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer"
g++ -O1 -g -std=c++17 -fPIC $SAN -I${CUDA_HOME:-/usr/local/cuda}/include -c $ROOT/tests/native/fake_cudart.cpp -o fake_cudart.o
g++ -shared $SAN -o liblp_fake.so $OUT/build/*.o fake_cudart.o
g++ -O1 -g -std=c++17 $SAN -I$ROOT/include $ROOT/tests/native/host_transform_fuzz.cpp -o transform_fake -L. -llp_fake -Wl,-rpath,$OUT/fake
g++ -O1 -g -std=c++17 $SAN -I$ROOT/include -I$ROOT/lilliput_b200/csrc -I${CUDA_HOME:-/usr/local/cuda}/include \
    $ROOT/tests/native/host_batch_fake_gpu.cpp -o batch_fake -L. -llp_fake -Wl,-rpath,$OUT/fake
# detect_leaks=0: the resize tap tables are cached for the life of the process
export ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1:max_allocation_size_mb=16384
./transform_fake $ITERS $OUT/seeds/* 2>&1 | grep -v "iterations$" | grep -v "^\[lilliput" | tail -5
g++ -O1 -g -std=c++17 $SAN -I$ROOT/include $ROOT/tests/native/host_abi_misuse_fake_gpu.cpp -o misuse_fake -L. -llp_fake -Wl,-rpath,$OUT/fake
./misuse_fake $((ITERS * 100)) 3 $OUT/seeds/* 2>&1 | grep -v "calls$" | grep -v "^\[lilliput" | grep -v "^Error: Final encoded" | tail -3
./batch_fake $((ITERS / 100 + 1)) $OUT/seeds/*jpeg_3* $OUT/seeds/*jpegvar* $OUT/seeds/*c1_input 2>&1 | grep -v "^\[lilliput" | tail -5
