#!/bin/bash
# usage: tools/gpu_launch_metrics.sh <tag> "<workloads>"   -- per-launch ncu metrics of the traversal kernels of one warm frame
#        -> gpurun_out/<tag>_lm_<wl>.csv and a table on stdout (tools/launch_metrics_table.py)
tag=$1; wls=${2:-"cfg3_noaa cfg4"}
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
for wl in $wls; do
    timeout 600 ncu --metrics $M --clock-control none -k 'regex:k_primary|k_closest|k_shadow' --csv --log-file gpurun_out/${tag}_lm_${wl}.csv \
        python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-other-workloads > /dev/null 2> gpurun_out/${tag}_lm_${wl}.err
    echo "== $wl"
    python tools/launch_metrics_table.py gpurun_out/${tag}_lm_${wl}.csv
done
