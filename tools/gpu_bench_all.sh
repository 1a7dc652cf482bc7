#!/bin/bash
# usage: tools/gpu_bench_all.sh <tag> "<workloads>" [lib]   -- bench lines of the named workloads into gpurun_out/<tag>_<wl>.json + a summary line each
tag=$1; wls=${2:-"cfg1 cfg2 cfg3_noaa cfg4"}; lib=$3
mkdir -p gpurun_out
for wl in $wls; do
    if [ -n "$lib" ]; then export PVGPU_LIB=$PWD/$lib; fi
    python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_${wl}.json 2> gpurun_out/${tag}_${wl}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${wl}.json"))
    k = d["roofline"]["kernel_ms_per_step"]
    print("${tag} ${wl}: %.3f ms/frame e2e %.3f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) + " ".join("%s %.2f" % (a, b) for a, b in k.items()))
except Exception as e:
    print("${tag} ${wl}: FAILED", e, open("gpurun_out/${tag}_${wl}.err").read()[-800:])
PY
done
