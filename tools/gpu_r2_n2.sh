#!/bin/bash
# round 2, two GPUs: NCCL slab parity test, the bench line at N=2 (slab_split included), kept logs for profiles/
TAG=${1:-r2n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_slab_gpu.py -q -m gpu --tb=short -p no:cacheprovider -s > $OUT/pytest_slab_gpu.log 2>&1; echo "slab pytest rc=$?"; tail -4 $OUT/pytest_slab_gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?"; tail -3 $OUT/bench_n2.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","slab_split")})
PY
