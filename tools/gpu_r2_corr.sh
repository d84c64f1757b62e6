#!/bin/bash
# round 2: correlate stage -- kernel parity tests, stage timing (two-kernel default, fused opt-in), bench line
TAG=${1:-r2c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_pipeline_gpu.py -q -m gpu --tb=short -p no:cacheprovider -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/time_corr.py 5 --exact 2>&1 | tail -1 | tee -a $OUT/time_corr.log
CA_CORR_FUSED=1 timeout 300 python tools/time_corr.py 5 2>&1 | tail -1 | tee -a $OUT/time_corr.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","stages_ms")}, d.get("roofline_correlate"))
for k in d.get("roofline_other_kernels", []) + [d["roofline"]]: print(k["kernel"][:40], k["avg_launch_ms"], k["launches_per_step"], round(k["frac"],3))
PY
tail -3 $OUT/bench.err
