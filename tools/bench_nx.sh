#!/bin/bash
# usage: tools/bench_nx.sh <tag> <ngpus> "<ENV=VAL ...>" <extra bench args...>
tag=$1; n=$2; envs=$3; shift 3
env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | tee gpurun_out/bench_n${n}_$tag.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$tag [$envs $*] N=$n value=%.0f lb=%.0f (%.3f ms) mp=%.0f (%.3f ms)'%(d['value'],d['lb']['mlups'],d['lb']['ms_per_step'],d['mp']['mlups'],d['mp']['ms_per_step']))"
