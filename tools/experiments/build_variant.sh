#!/bin/bash
# Build a library variant into build_variants/<name>.so without touching the tree's own build:
#   tools/experiments/build_variant.sh w10c2 "-DOPN_FRAME_WARPS=10 -DOPN_FRAME_CTAS=2" ["-DOPN_FRAME_GROUPS=1"]
# $2 = extra nvcc flags (kernels), $3 = extra g++ flags (host runtime)
set -e
name=$1; nv=$2; cx=$3
root=$(cd "$(dirname "$0")/../.." && pwd)
tmp=$(mktemp -d)
cp -r $root/opus-native_b200/csrc $tmp/csrc
mkdir -p $tmp/include && cp $root/include/*.h $tmp/include/
mkdir -p $tmp/x/y && mv $tmp/csrc $tmp/x/y/csrc && mkdir -p $tmp/include && mv $tmp/include $tmp/x/include 2>/dev/null || true
# csrc includes ../../include/opusb200.h
cd $tmp/x/y/csrc && rm -rf build && make -s EXTRA="$nv" CXXEXTRA="$cx" OUT=$tmp/lib.so >/dev/null
mkdir -p $root/build_variants && cp $tmp/lib.so $root/build_variants/$name.so
rm -rf $tmp
echo "built build_variants/$name.so"
