#!/bin/bash
# quick GPU session: parity tests, stage timings, short bench.  usage (under gpurun): bash tools/gpu_quick.sh <tag> [pytest-args]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
shift
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x "$@" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
timeout 300 python tools/stage_times.py 5 2>&1 | tail -1 | tee -a $OUT/stage_times.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
