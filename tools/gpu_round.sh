#!/bin/bash
# One GPU session: smoke, parity tests, bench, ncu launch list + full captures.  Everything lands in gpurun_out/.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/host.txt; free -g >> $OUT/host.txt
echo "== smoke" ; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -25 $OUT/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
cat $OUT/bench.json; tail -5 $OUT/bench.err
if [ "$2" != "quick" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python tools/profile_pair.py 10 > $OUT/ncu_list.log 2>&1; echo "ncu list rc=$?"
tail -2 $OUT/ncu_list.log
echo "== ncu full captures (dominant kernels; the full set is tools/gpu_profiles_r1.sh)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box12_fast -s 1 -c 1 -f -o $OUT/prof_k_box12_fast python tools/profile_corr.py --fast > $OUT/ncu_full1.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_raw_ssd_staged -s 1 -c 1 -f -o $OUT/prof_k_raw_ssd_staged python tools/profile_corr.py --fast > $OUT/ncu_full2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam_grad -s 1 -c 1 -f -o $OUT/prof_k_adam_grad python tools/profile_pair.py 2 > $OUT/ncu_full3.log 2>&1; echo "rc=$?"
ls -la $OUT
fi
echo "== memcheck (small)"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log
tail -4 $OUT/memcheck.log
