#!/bin/bash
# Experiment builds of the library: tools/build_variants.sh name "flags" [name "flags" ...]  ->  variants/libpvgpu_<name>.so
# (picked up with PVGPU_LIB=...; variants/ is git-ignored but travels to the GPU box)
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    make -s -j8 LIB=variants/libpvgpu_$name.so OBJDIR=build/obj_$name EXTRA="$flags" 2>&1 | grep -E "error|Error" -A3 || true
    ls -la variants/libpvgpu_$name.so
done
