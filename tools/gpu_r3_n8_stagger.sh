mkdir -p gpurun_out/r3n8b
for st in 0 1; do
BENCH_E2E_STAGGER=$st timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$st bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r3n8b/bench_n8_stagger$st.json 2> gpurun_out/r3n8b/err$st.log
python - <<PY
import json
d=json.loads(open("gpurun_out/r3n8b/bench_n8_stagger$st.json").read().strip().splitlines()[-1])
print("stagger $st", {k:d.get(k) for k in ("value","ms_per_step","e2e","e2e_pipelined")})
PY
done
