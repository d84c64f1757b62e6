#!/bin/bash
# tuning session: parity tests first, then stage timings under the library's tuning knobs
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
for minb in 6 4 8; do for chunk in 1 2 4; do
  CA_RAW_MINB=$minb CA_CORR_CHUNK=$chunk timeout 300 python tools/stage_times.py 5 2>&1 | tail -1 | tee -a $OUT/stage_times.log
done; done
CA_CORR_CHUNK=13 timeout 300 python tools/stage_times.py 5 2>&1 | tail -1 | tee -a $OUT/stage_times.log
timeout 300 python tools/stage_times.py 5 --e2e 2>&1 | tail -9 | tee -a $OUT/e2e.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
