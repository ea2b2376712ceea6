#!/bin/bash
# usage: tools/ncu_one.sh <name> <kernel regex> <workload> <nlb> <nmp> [ENV=VAL ...]  -- one ncu --set full capture (last launches)
name=$1; regex=$2; wl=$3; nlb=$4; nmp=$5; shift 5
mkdir -p gpurun_out
env "$@" ncu --set full --clock-control none --import-source on -k regex:"$regex" -s 4 -c 1 -f -o gpurun_out/prof_$name python tools/profile_run.py $wl $nlb $nmp > gpurun_out/ncu_$name.log 2>&1
tail -n 2 gpurun_out/ncu_$name.log
