#!/bin/bash
# round 2: GPU parity tests with the achieved errors printed (-s), kept under gpurun_out/<tag>/
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^\[(cuda|oracle|adam)" $OUT/pytest_gpu.log | sort | uniq > $OUT/parity_achieved.txt
tail -5 $OUT/pytest_gpu.log; cat $OUT/parity_achieved.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json; tail -3 $OUT/bench.err
