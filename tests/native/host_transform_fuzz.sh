#!/bin/bash
# See host_transform_fuzz.cpp.  Needs oracle/_ref/*.o (make -C oracle ref) and /root/reference's vendored libs.
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
R=${R:-/root/reference}
D=$R/deps/linux/amd64
OUT=${LP_ASAN_DIR:-/tmp/asan}
ITERS=${1:-20000}
mkdir -p $OUT/seeds
[ -n "$(ls $OUT/seeds 2>/dev/null)" ] || { echo "run tests/native/host_parse_fuzz.sh once first (it dumps the seed files)"; exit 2; }
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -fno-sanitize-recover=undefined"
g++ -std=c++20 -O1 -g -fPIC -shared $SAN -DLP_REFERENCE_BACKEND -I$ROOT/include -I$ROOT/lilliput_b200/host \
  -I$D/include/opencv4 -o $OUT/libref_asan.so $ROOT/oracle/ref_shim.cpp $ROOT/lilliput_b200/host/lilliput_host.cpp \
  $ROOT/oracle/_ref/opencv.o $ROOT/oracle/_ref/webp.o $ROOT/oracle/_ref/giflib.o \
  -Wl,-Bsymbolic -Wl,--exclude-libs,ALL -L$D/lib -L$D/lib/opencv4/3rdparty \
  -lopencv_photo -lopencv_imgcodecs -lopencv_imgproc -lopencv_core -lgif -ljpeg -lpng16 \
  -lwebpmux -lwebpdemux -lwebp -lsharpyuv -lz -llibopenjp2 -littnotify -lippiw -lippicv -llcms2 -lpthread -ldl
g++ -O1 -g -std=c++17 $SAN -I$ROOT/include $ROOT/tests/native/host_transform_fuzz.cpp -o $OUT/host_transform_fuzz \
  -L$OUT -lref_asan -Wl,-rpath,$OUT
# The vendored OpenCV reads freed EXIF marker memory on some mutated JPEGs (cv::ExifReader::parseExif under
# cv::JpegDecoder::readHeader): the reference's code, not the layer under test -- suppressed by function name.
printf 'interceptor_via_fun:parseExif\ninterceptor_via_fun:cv::ExifReader::parseExif\n' > $OUT/asan.supp
# detect_leaks=0: the vendored PngDecoder leaks its chunk buffers when libpng longjmps out of a broken file
ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1:max_allocation_size_mb=4096:suppressions=$OUT/asan.supp $OUT/host_transform_fuzz $ITERS $OUT/seeds/*
