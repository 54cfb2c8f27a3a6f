#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu --maxfail=12 -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
timeout 500 python tools/train_e2e.py 32 > gpurun_out/train_e2e_ring.json 2> gpurun_out/train_e2e_ring.err; echo "e2e rc=$?"; cat gpurun_out/train_e2e_ring.json
C2V_BATCH_RING=0 timeout 500 python tools/train_e2e.py 32 > gpurun_out/train_e2e_noring.json 2> gpurun_out/train_e2e_noring.err; cat gpurun_out/train_e2e_noring.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print(d["ms_per_step"], d["value"], d["e2e"], d["fp32_equivalent"], d["roofline"], d["gather_roofline"]); print({k:v["ms"] for k,v in d["phases"].items()}); print(d["cpu_baseline"])
PY
