#!/bin/bash
# round 2: the full bench line (extras included) and the reference arm, as the driver runs them
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","stages_ms","value_t0_t1","e2e","e2e_pipelined","cpu_baseline","gpu_launches","clocks")})
print("roofline", {k:d["roofline"][k] for k in ("kernel","achieved","frac","avg_launch_ms","traffic")})
print("other_configs", d.get("other_configs")); print("reference_torch_cuda", d.get("reference_torch_cuda"))
PY
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cat $OUT/bench_ref.json | cut -c1-900
