#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_corr.csv python tools/profile_corr.py --cc > $OUT/ncu_list.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box12 -s 1 -c 1 -f -o $OUT/prof_box12 python tools/profile_corr.py > $OUT/ncu_full1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raw_ssd -s 1 -c 1 -f -o $OUT/prof_raw_ssd python tools/profile_corr.py > $OUT/ncu_full2.log 2>&1; echo "rc=$?"
ls -la $OUT
