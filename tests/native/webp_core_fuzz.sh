#!/bin/bash
# See webp_core_fuzz.cpp.  CPU only; seeds = the WebP golden fixtures (dumped by host_parse_fuzz.sh).
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
OUT=${LP_ASAN_DIR:-/tmp/asan}
ITERS=${1:-200000}
[ -n "$(ls $OUT/seeds 2>/dev/null)" ] || { echo "run tests/native/host_parse_fuzz.sh once first (it dumps the seed files)"; exit 2; }
# signed overflow / shifts of negative values on garbage coefficients wrap on the device exactly as here (and as in
# libwebp's own MUL macros); the probe is after memory errors
g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize=signed-integer-overflow,shift-base -fno-omit-frame-pointer -fno-sanitize-recover=undefined \
    $ROOT/tests/native/webp_core_fuzz.cpp -o $OUT/webp_core_fuzz
ASAN_OPTIONS=detect_leaks=1:allocator_may_return_null=1 $OUT/webp_core_fuzz $ITERS $OUT/seeds/*webp_*
