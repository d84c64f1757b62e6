#!/bin/bash
# round-2 evidence: launch list of the bench command + one `ncu --set full` capture per main kernel
# (read back here with tools/ncu_summary.py / tools/ncu_regions.py).  gpurun returns <= 64 MiB per call: two parts.
OUT=gpurun_out/${1:-r2prof}
PART=${2:-all}
mkdir -p $OUT
if [ "$PART" = "all" ] || [ "$PART" = "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_under_ncu.log 2>&1; echo "list rc=$?"
for k in k_box12_fast k_raw_ssd_tmap k_blockmin_argmin; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 0 -c 1 -f -o $OUT/prof_$k python tools/profile_corr.py --fast --cc > /dev/null 2>&1; echo "$k rc=$?"
done
CA_CORR_FUSED=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_corr_fused -s 0 -c 1 -f -o $OUT/prof_k_corr_fused python tools/profile_corr.py --fast > /dev/null 2>&1; echo "fused rc=$?"
fi
if [ "$PART" = "all" ] || [ "$PART" = "2" ]; then
for k in k_adam_grad k_box3_h k_box3_p; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/prof_$k python tools/profile_pair.py 2 > /dev/null 2>&1; echo "$k rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python tools/dbg_fused.py 12 9 30 70 6 > $OUT/sanitizer_fused_memcheck.log 2>&1; tail -3 $OUT/sanitizer_fused_memcheck.log
CA_CORR_FUSED=1 timeout 900 compute-sanitizer --tool racecheck python tools/dbg_fused.py 12 5 14 40 3 > $OUT/sanitizer_fused_racecheck.log 2>&1; tail -3 $OUT/sanitizer_fused_racecheck.log
fi
ls -la $OUT
