#!/bin/bash
# round 2: the fused correlate kernel -- kernel parity tests, stage timing fused vs two-kernel, one ncu capture
TAG=${1:-r2f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
CA_CORR_FUSED=0 timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
if [ "$2" != "noncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_corr_fused -c 1 -o $OUT/prof_k_corr_fused -f python tools/profile_corr.py --fast > $OUT/ncu.log 2>&1; tail -2 $OUT/ncu.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","stages_ms")}, d.get("roofline_correlate"))
PY
tail -3 $OUT/bench.err
