#!/bin/bash
# first GPU pass of round 2: tests, then the bench lines of every mode
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi0.csv 2>&1
timeout 900 python -m pytest tests -m gpu --maxfail=12 -q -p no:cacheprovider 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default rc=$?"; cut -c1-600 gpurun_out/bench_default.json
timeout 200 python bench.py --steps 20 --warmup 5 --math 3xtf32 --no-cpu-baseline > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench_3xtf32.err
echo "bench 3x rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --mode fwd_loss --no-cpu-baseline > gpurun_out/bench_fwd_loss.json 2> gpurun_out/bench_fwd_loss.err
echo "bench fwd_loss rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --mode sampled --no-cpu-baseline > gpurun_out/bench_sampled.json 2> gpurun_out/bench_sampled.err
echo "bench sampled rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-lazy-adam --no-cpu-baseline --no-fp32-equivalent > gpurun_out/bench_dense_adam.json 2> gpurun_out/bench_dense_adam.err
echo "bench dense rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --zipf --bags normal --no-cpu-baseline --no-fp32-equivalent > gpurun_out/bench_zipf.json 2> gpurun_out/bench_zipf.err
echo "bench zipf rc=$?"
for f in 3xtf32 fwd_loss sampled dense_adam zipf; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$f.json"))
    print("$f", d["ms_per_step"], d["value"], d["e2e"]["value"], {k:v["ms"] for k,v in d["phases"].items()})
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/bench_$f.err").read()[-1500:])
PY
done
