#!/bin/bash
# usage: tools/bench_n2.sh <tag> "<ENV=VAL ...>" <extra bench args...>
tag=$1; envs=$2; shift 2
env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$tag [$envs $*] N=2 value=%.0f lb=%.0f (%.3f ms) mp=%.0f (%.3f ms)'%(d['value'],d['lb']['mlups'],d['lb']['ms_per_step'],d['mp']['mlups'],d['mp']['ms_per_step']))"
