#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_adam|k_field_box|k_avgpool|k_planar' --csv --log-file $OUT/launches_adam.csv python tools/profile_pair.py 3 > $OUT/ncu_list_adam.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam_grad -s 1 -c 1 -f -o $OUT/prof_adam_grad python tools/profile_pair.py 2 > $OUT/ncu_full_adam.log 2>&1; echo "rc=$?"
