#!/bin/bash
# Experiment builds of the library: tools/build_variants.sh name "flags" [name "flags" ...]  ->  variants/libpvgpu_<name>.so
# (picked up with PVGPU_LIB=...; variants/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    # compile from a snapshot of the sources, so that the working tree can be edited while variants build
    rm -rf build/src_$name && mkdir -p build/src_$name/include && cp -r povray_b200/csrc build/src_$name/csrc && cp include/pvgpu.h build/src_$name/include/
    make -s -j${JOBS:-8} LIB=variants/libpvgpu_$name.so OBJDIR=build/obj_$name CSRC=build/src_$name/csrc INCDIR=build/src_$name/include EXTRA="$flags" 2>&1 | grep -E "error|Error" -A3 || true
    ls -la variants/libpvgpu_$name.so
done
