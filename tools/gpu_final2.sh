#!/bin/bash
# Final single-GPU pass of round 2: GPU suite, the default bench line (with the CPU arm), the other bench modes,
# one `ncu --set full` capture of a step and the launch list of the same command.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/f2_default.json 2> gpurun_out/f2_default.err || tail -c 600 gpurun_out/f2_default.err
timeout 60 python bench.py --steps 20 --warmup 5 --mode fwd_loss --no-cpu-baseline > gpurun_out/f2_fwd_loss.json 2>/dev/null
timeout 60 python bench.py --steps 20 --warmup 5 --mode sampled --no-cpu-baseline > gpurun_out/f2_sampled.json 2>/dev/null
timeout 60 python bench.py --steps 20 --warmup 5 --zipf --bags normal --no-cpu-baseline --no-fp32-equivalent > gpurun_out/f2_zipf.json 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -s 120 -c 34 -o gpurun_out/prof_step2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-equivalent > gpurun_out/ncu_step2.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 70 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-fp32-equivalent > /dev/null 2>&1
python - <<'PY'
import json
for n in ("default", "fwd_loss", "sampled", "zipf"):
    try:
        d = json.loads([l for l in open("gpurun_out/f2_%s.json" % n) if l.startswith("{")][-1])
        fe = d.get("fp32_equivalent") or {}
        print(n, d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "3x", fe.get("ms_per_step"), fe.get("value"), "roof", d["roofline"]["kernel"], d["roofline"]["frac"],
              "gather", (d.get("gather_roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        if n == "default": print("   ", {k: v["ms"] for k, v in d["phases"].items()})
    except Exception as ex:
        print(n, "ERR", ex)
PY
ls -la gpurun_out/prof_step2.ncu-rep gpurun_out/launches_r02b.csv
