#!/bin/bash
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_adam|k_box3|k_avgpool|k_planar' --csv --log-file $OUT/launches_adam.csv python tools/profile_pair.py 3 > $OUT/ncu_list_adam.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_adam_grad -s 1 -c 1 -f -o $OUT/prof_adam_grad python tools/profile_pair.py 2 > $OUT/ncu_full_adam.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box3_h -s 1 -c 1 -f -o $OUT/prof_box3_h python tools/profile_pair.py 2 > $OUT/ncu_full_adam2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box3_p -s 1 -c 1 -f -o $OUT/prof_box3_p python tools/profile_pair.py 2 > $OUT/ncu_full_adam3.log 2>&1; echo "rc=$?"
python - <<'PY'
import sys, time, os
sys.path.insert(0, os.getcwd())
import torch
from convexadam_b200.convex_adam_MIND import convex_adam_pt
from convexadam_b200.synth import make_pair
dev = torch.device("cuda:0")
f, m = make_pair((192, 160, 256), seed=0)
fp, mp = f.pin_memory(), m.pin_memory()
kw = dict(mind_r=1, mind_d=2, lambda_weight=1.25, grid_sp=2, disp_hw=6, selected_niter=100, selected_smooth=0, grid_sp_adam=2, ic=True, dtype=torch.float32, device=dev)
for i in range(8):
    t0 = time.perf_counter(); out = convex_adam_pt(fp, mp, **kw); t1 = time.perf_counter()
    print(f"e2e call {i}: {1e3*(t1-t0):.2f} ms (keeping the previous result alive)")
for i in range(4):
    t0 = time.perf_counter(); convex_adam_pt(fp, mp, **kw); t1 = time.perf_counter()
    print(f"e2e call {i}: {1e3*(t1-t0):.2f} ms (result dropped)")
PY
