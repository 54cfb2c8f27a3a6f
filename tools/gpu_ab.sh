#!/bin/bash
# A/B of the round's late single-GPU changes: $1.. = extra bench flag sets, one run each (after the default run)
show() { python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        fe = d.get("fp32_equivalent") or {}
        print(f, d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "3x", fe.get("ms_per_step"), "fb", d["config"].get("exp_slab_fallbacks"), "loss", d["config"]["last_loss"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"])
        print("   ", {k: v["ms"] for k, v in d["phases"].items()})
    except Exception as ex:
        print(f, "ERR", ex)
PY
}
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err || tail -c 800 gpurun_out/ab_default.err
show gpurun_out/ab_default.json
i=0
for flags in "$@"; do
  i=$((i+1))
  timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-fp32-equivalent $flags > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -c 800 gpurun_out/ab_$i.err
  echo "== $flags"; show gpurun_out/ab_$i.json
done
