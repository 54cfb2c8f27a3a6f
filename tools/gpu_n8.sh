#!/bin/bash
# 8-GPU pass: java14m weak scaling line and BASELINE configs[4] ("large": 3M/2M vocab, d=256, target table row-sharded),
# each with the push-based gradient exchange (default) and with remote red.add
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${1:-8}
run() { name=$1; shift; nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before_$name.txt 2>&1
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 100)) bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_$name.txt 2>&1
  python - <<PY
import json
try:
    txt=[l for l in open("gpurun_out/$name.json") if l.startswith("{")][0]
    d=json.loads(txt)
    print("$name", d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"]["last_loss"], d["roofline"]["kernel"], d["roofline"]["frac"])
    print("   ", {k:(v["ms"], v.get("gbs")) for k,v in d["phases"].items()})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/$name.err").read()[-1500:])
PY
}
run bench_n${N} --steps 20 --warmup 5
run bench_n${N}_nopush --steps 20 --warmup 5 --no-push-grads
run bench_large_n${N} --steps 10 --warmup 3 --workload large
run bench_large_n${N}_nopush --steps 10 --warmup 3 --workload large --no-push-grads
