#!/bin/bash
# usage: tools/bench_libs.sh <tag> "<lib suffixes, e.g. '- _exp1 _exp2'>" <workloads...>  -- device-only bench numbers per lib variant
tag=$1; libs=$2; shift 2
for lib in $libs; do
  [ "$lib" = "-" ] && lib=""
  for wl in "$@"; do
    f=laboetie_b200/lib/liblaboetie_gpu$lib.so
    [ -f $f ] || continue
    LBG_LIB=$PWD/$f python bench.py --steps 30 --warmup 3 --workload $wl --no-cpu-baseline --no-e2e --also "" $LBG_BENCH_ARGS 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$tag lib=$lib wl=$wl value=%.0f lb=%.0f (%.3f ms, frac %.3f) mp=%.0f (%.3f ms, frac %.3f)'%(d['value'],d['lb']['mlups'],d['lb']['ms_per_step'],d['roofline']['frac'],d['mp']['mlups'],d['mp']['ms_per_step'],d['roofline']['mp_step_kernel']['frac']))"
  done
done
