#!/bin/bash
# usage: tools/ab.sh <tag> <workload> <steps> "<lib suffix or ->|ENV=VAL ENV=VAL" ...   -- A/B of library builds and tuning variables on one box
tag=$1; wl=$2; steps=$3; shift 3
for spec in "$@"; do
  lib=${spec%%|*}; envs=${spec#*|}
  [ "$lib" = "-" ] && lib=""
  f=$PWD/laboetie_b200/lib/liblaboetie_gpu$lib.so
  [ -f $f ] || { echo "missing $f"; continue; }
  env LBG_LIB=$f $envs python bench.py --steps $steps --warmup 3 --workload $wl --no-cpu-baseline --no-e2e --also "" 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read())
    print('$tag lib=$lib [$envs] wl=$wl value=%.0f lb %.3f ms frac %.3f | mp %.3f ms frac %.3f'%(d['value'],d['lb']['ms_per_step'],d['roofline']['frac'],d['mp']['ms_per_step'],d['roofline']['mp_step_kernel']['frac']))
except Exception as e:
    print('$tag lib=$lib [$envs] FAILED', e)"
done
