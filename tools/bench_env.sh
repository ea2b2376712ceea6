#!/bin/bash
# usage: tools/bench_env.sh <tag> "<ENV=VAL ...>" <workloads...>  -- device-only bench numbers under an env setting
tag=$1; envs=$2; shift 2
for wl in "$@"; do
  env $envs python bench.py --steps 30 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e --also "" $LBG_BENCH_ARGS 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$tag [$envs] wl=$wl value=%.0f lb=%.0f (%.3f ms, frac %.3f) mp=%.0f (%.3f ms, frac %.3f)'%(d['value'],d['lb']['mlups'],d['lb']['ms_per_step'],d['roofline']['frac'],d['mp']['mlups'],d['mp']['ms_per_step'],d['roofline']['mp_step_kernel']['frac']))"
done
