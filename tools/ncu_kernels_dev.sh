#!/bin/bash
# developer helper (GPU box): per-kernel ncu durations / instruction counts of a few resident cycles
#   tools/ncu_kernels_dev.sh VARIANT DISTRIBUTION OUT.csv
v=$1; d=$2; out=$3
if [ "$v" = main ]; then unset KOMPASS_B200_LIB; else export KOMPASS_B200_LIB=$PWD/kompass-core_b200/lib/variants/libkompass_b200_$v.so; fi
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --launch-skip 60 -c 44 --csv --log-file $out python tools/family_dev.py --replay $d 12 > /dev/null 2>&1
