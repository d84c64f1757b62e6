#!/bin/bash
# round 2: the producer / consumer correlate kernel (correlate_pipe.cu) -- parity tests, stage timing against the two-kernel path,
# knob sweep.  Every step under its own timeout: the kernel's CTAs wait for each other.
TAG=${1:-r2p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 120 python tools/dbg_pipe.py 12 9 30 64 6 2>&1 | tail -3 | tee $OUT/dbg_small.log
timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
CA_CORR_PIPE=0 timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
for nw in 3 2; do CA_PIPE_NW=$nw timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log; done
for r in 8 32; do CA_PIPE_R=$r timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log; done
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x -k "certified or fast or full_size" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
