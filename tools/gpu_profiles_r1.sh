#!/bin/bash
# round-1 evidence: launch list of the bench command + one full capture per main kernel (read back with tools/ncu_summary.py)
# NB gpurun brings back at most 64 MiB per call: the ten captures are ~50 MB, run the two loops in separate calls if it refuses
OUT=gpurun_out/${1:-r1final}
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_under_ncu.log 2>&1; echo "list rc=$?"
for k in k_box12_fast k_raw_ssd_staged k_blockmin_argmin k_cost_argmin_sep k_cc_fixup; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o $OUT/prof_$k python tools/profile_corr.py --fast --cc > /dev/null 2>&1; echo "$k rc=$?"
done
for k in k_adam_grad k_box3_h k_box3_p k_mind_phase_a k_mind_phase_b; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/prof_$k python tools/profile_pair.py 2 > /dev/null 2>&1; echo "$k rc=$?"
done
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; tail -c 600 $OUT/bench.json
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err; tail -c 400 $OUT/bench_reference.json
ls $OUT
