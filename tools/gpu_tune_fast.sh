#!/bin/bash
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/time_corr.py 5 --exact 2>&1 | tail -1 | tee -a $OUT/time_corr.log
for minb in 4 5 6 8; do
  CA_RAW_MINB=$minb timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
done
for bt in 256 320 384; do
  CA_BOX_THREADS=$bt timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
done
for chunk in 8 20; do
  CA_CORR_CHUNK=$chunk timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
done
