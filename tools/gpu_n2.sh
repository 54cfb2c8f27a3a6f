#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out

run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 100)) bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python - <<PY
import json
try:
    txt=[l for l in open("gpurun_out/$name.json") if l.startswith("{")][0]
    d=json.loads(txt)
    print("$name", d["ms_per_step"], d["value"], d["config"]["last_loss"], {k:(v["ms"], v.get("gbs")) for k,v in d["phases"].items() if k in ("gather","dx_scatter","dW","adam","peer_sort","inbox_apply")})
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/$name.err").read()[-1500:])
PY
}
run n2_java_push
run n2_java_nopush --no-push-grads


