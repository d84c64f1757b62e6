#!/bin/bash
# round 2, third session: parity tests, the full bench line + reference arm, launch list, ncu captures of the kernels
# this session changed (Adam gradient second form, quad trilinear, inverse consistency), memcheck over their tests
TAG=${1:-r3e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^\[(cuda|oracle|adam)|^\.*\[(cuda|oracle|adam)" $OUT/pytest_gpu.log | sed 's/^\.*//' | sort | uniq > $OUT/parity_achieved.txt
tail -3 $OUT/pytest_gpu.log
timeout 1200 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","stages_ms","value_t0_t1","value_two_stream_directions","e2e","e2e_pipelined","gpu_launches","clocks")})
print("roofline", {k:d["roofline"][k] for k in ("kernel","achieved","frac","avg_launch_ms","traffic")})
print("other_configs", d.get("other_configs"))
PY
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; cut -c1-400 $OUT/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > $OUT/bench_under_ncu.log 2>&1; echo "list rc=$?"
for k in k_adam_grad2 k_trilinear_q k_ic_step; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $OUT/prof_$k python tools/profile_pair.py 2 > /dev/null 2>&1; echo "$k rc=$?"
done
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "adam or trilinear or inverse or consistency" > $OUT/sanitizer_memcheck.log 2>&1; tail -4 $OUT/sanitizer_memcheck.log
ls -la $OUT
