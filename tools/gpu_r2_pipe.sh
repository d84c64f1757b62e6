#!/bin/bash
# round 2: the producer / consumer correlate kernel (correlate_pipe.cu, opt-in) -- parity tests, stage timing against the
# two-kernel path, one ncu capture.  Every step under its own timeout: the kernel's CTAs wait for each other.
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
CA_CORR_PIPE=1 timeout 120 python tools/dbg_pipe.py 12 9 30 64 6 2>&1 | tail -3 | tee $OUT/dbg_small.log
CA_CORR_PIPE=1 timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x -k "pipe" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
CA_CORR_PIPE=1 timeout 800 ncu --set full --clock-control none --import-source on -k regex:k_corr_pipe -c 1 -f -o $OUT/prof_k_corr_pipe python tools/profile_corr.py --fast > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
