#!/bin/bash
# usage: tools/run_scale.sh <tag> <ngpus> [what...]  -- weak (cfg5w, with e2e) and strong (cfg4, cfg5s) scaling lines on one box
tag=$1; n=$2; shift 2; what=${*:-"cfg5w cfg5s cfg4 cfg5s_aa"}
mkdir -p gpurun_out
run() { # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + RANDOM % 200)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${name}_n${n}_$tag.out 2> gpurun_out/${name}_n${n}_$tag.err
  tail -1 gpurun_out/${name}_n${n}_$tag.out > gpurun_out/${name}_n${n}_$tag.json
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${name}_n${n}_$tag.json').read())
    e=d.get('e2e') or {}
    print('${name} N=$n', round(d['value']), 'MLUPS lb %.3f ms mp %.3f ms'%(d['lb']['ms_per_step'], d['mp']['ms_per_step']), d['scaling'], d['config']['phase_a_layout'], 'verify', d['verify'].get('ok'), d['verify'].get('path'), 'e2e', round(e.get('value',0)), {k:round(v,3) for k,v in (e.get('phase_seconds') or {}).items()})
except Exception as ex:
    print('${name} N=$n FAILED', ex); print(open('gpurun_out/${name}_n${n}_$tag.err').read()[-1200:])
PY
}
for w in $what; do
  case $w in
    cfg5w) run cfg5w ;;
    cfg5w_aa) run cfg5w_aa --no-e2e --in-place ;;
    cfg5s) run cfg5s --workload cfg5s --no-e2e ;;
    cfg4) run cfg4 --workload cfg4 --no-e2e ;;
    cfg5s_aa) run cfg5s_aa --workload cfg5s --no-e2e --in-place ;;
  esac
done
