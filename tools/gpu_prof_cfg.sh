#!/bin/bash
# usage: tools/gpu_prof_cfg.sh <workload> <tag> [kernel-regex]   -- one ncu --set full capture of the first launches of the traversal kernels
wl=$1; tag=$2; rx=${3:-"k_closest|k_shadow"}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 4 -c 2 -o gpurun_out/prof_${tag} -f \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_${tag}.log 2>&1
tail -3 gpurun_out/prof_${tag}.log
