#!/bin/bash
# 8-GPU pass: java14m weak scaling line and BASELINE configs[4] ("large": 3M/2M vocab, d=256, target table row-sharded)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before.txt 2>&1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
echo "java14m N=$N rc=$?"
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_java14m.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --workload large > gpurun_out/bench_large_n${N}.json 2> gpurun_out/bench_large_n${N}.err
echo "large N=$N rc=$?"
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_large.txt 2>&1
for f in bench_n${N} bench_large_n${N}; do python - <<PY
import json
try:
    txt=[l for l in open("gpurun_out/$f.json") if l.startswith("{")][0]
    d=json.loads(txt)
    print("$f", d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"])
    print({k:(v["ms"], v.get("gbs"), v.get("tflops")) for k,v in d["phases"].items()})
except Exception as e:
    print("$f FAILED", e); print(open("gpurun_out/$f.err").read()[-2500:])
PY
done
head -12 gpurun_out/nvlink_after_large.txt
