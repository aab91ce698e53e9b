#!/bin/bash
# developer (GPU box): the round's measurement pass - launch lists of the cycle per distribution and of one
# sweep chunk set, one `ncu --set full` capture of the cycle kernels, then the bench itself (never under ncu)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
set -x
python tools/family_dev.py --replay dense_cluster_on_path 12 > gpurun_out/plain1.log 2>&1 || exit 1
for d in dense_cluster_on_path friendly_ring; do
  ncu --metrics $M --clock-control none --launch-skip 60 -c 160 --csv --log-file gpurun_out/r2_launches_$d.csv python tools/family_dev.py --replay $d 12 > /dev/null 2>&1
done
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_sweep.csv python tools/sweep_prof_dev.py dense_cluster_on_path 64 > /dev/null 2>&1
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_sweep_friendly_ring.csv python tools/sweep_prof_dev.py friendly_ring 64 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 24 -c 24 -o gpurun_out/r2_cycle_full python tools/family_dev.py --replay dense_cluster_on_path 12 > gpurun_out/ncu_full.log 2>&1
# (bench.py reads the instruction counts of profiles/r2_cycle_ncu_summary.json: digest the launch lists with
#  tools/ncu_family.py first, then run `python bench.py > gpurun_out/r2_bench_1gpu.log` in a second call)
