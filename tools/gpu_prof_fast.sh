#!/bin/bash
# launch list + full captures of the certified path's kernels (one direction at config 2)
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_fast.csv python tools/profile_corr.py --fast --cc > $OUT/ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box12 -s 1 -c 1 -f -o $OUT/prof_box12_fast python tools/profile_corr.py --fast > $OUT/ncu_full1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raw_ssd -s 1 -c 1 -f -o $OUT/prof_raw_fast python tools/profile_corr.py --fast > $OUT/ncu_full2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cost_argmin_pruned -s 2 -c 1 -f -o $OUT/prof_pruned_cert python tools/profile_corr.py --fast --cc > $OUT/ncu_full3.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cc_fixup -s 2 -c 1 -f -o $OUT/prof_fixup python tools/profile_corr.py --fast --cc > $OUT/ncu_full4.log 2>&1; echo "rc=$?"
ls -la $OUT
