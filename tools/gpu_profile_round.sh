#!/bin/bash
# GPU-box side of the per-round profile refresh: the launch list of one benchmark invocation and one ncu --set full capture of the first
# launches of the traversal kernels, for configs 2 and 3 (no AA).  usage: tools/gpu_profile_round.sh <round-tag>   (outputs in gpurun_out/)
tag=${1:-r2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/${tag}_launches_cfg2.log 2>&1
for wl in cfg2 cfg3_noaa; do
    ncu --set full --clock-control none --import-source on -k "regex:k_closest|k_shadow" -s 14 -c 4 -o gpurun_out/${tag}_prof_${wl} -f \
        python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > gpurun_out/${tag}_prof_${wl}.log 2>&1
    tail -2 gpurun_out/${tag}_prof_${wl}.log | cut -c1-200
done
