#!/bin/bash
# round 2, eight GPUs: the bench line at N=8 with slab_split (config-2 pair over 8 ranks + BASELINE config 5: 384^3, disp_hw=8)
TAG=${1:-r2n8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > $OUT/bench_n8.json 2> $OUT/bench_n8.err; echo "bench rc=$?"; tail -5 $OUT/bench_n8.err
python - <<PY
import json
d=json.loads(open("$OUT/bench_n8.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","e2e_pipelined","slab_split")})
PY
