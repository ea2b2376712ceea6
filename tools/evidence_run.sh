#!/bin/bash
# usage: tools/evidence_run.sh <tag>   -- one-GPU evidence pass: tests, smoke, both bench arms, ncu launch list + full capture
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py 2> gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_cfg5w_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_ref_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5w_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --also "" > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lb_step_kernel|mp_step_kernel" -s 4 -c 4 -f -o gpurun_out/prof_cfg5w_$tag \
    python tools/profile_run.py cfg5w 6 6 > gpurun_out/ncu_a.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/bench_cfg5w_$tag.json", "gpurun_out/bench_ref_$tag.json"):
    try:
        d = json.load(open(f))
        print(f, "value", d.get("value"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("frac"),
              "mp", ((d.get("roofline") or {}).get("mp_step_kernel") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
