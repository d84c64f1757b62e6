#!/bin/bash
# parity tests, stage timings and the bench line (no profiler)
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/time_corr.py 5 --exact 2>&1 | tail -1 | tee -a $OUT/time_corr.log
timeout 300 python tools/stage_times.py 5 --e2e 2>&1 | tail -10 | tee -a $OUT/stage_times.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
