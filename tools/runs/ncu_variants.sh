#!/bin/bash
# usage: tools/ncu_variants.sh <tag>  -- ncu --set full on the LB step kernel variants and the streams2 skeletons (one box)
tag=$1
mkdir -p gpurun_out
LBG_LB_TPC=0 LBG_MP_TPC=0 ncu --set full --clock-control none --import-source on -k regex:"lb_step_kernel" -s 4 -c 1 -f -o gpurun_out/prof_lb_static_$tag python tools/profile_run.py cfg5w 6 0 > gpurun_out/ncu_1.log 2>&1
LBG_LB_TPC=8 LBG_MP_TPC=0 ncu --set full --clock-control none --import-source on -k regex:"lb_step_kernel" -s 4 -c 1 -f -o gpurun_out/prof_lb_dyn8_$tag python tools/profile_run.py cfg5w 6 0 > gpurun_out/ncu_2.log 2>&1
LBG_LB_TPC=0 LBG_LB_MINB=0 LBG_MP_TPC=0 ncu --set full --clock-control none --import-source on -k regex:"lb_step_async_kernel" -s 4 -c 1 -f -o gpurun_out/prof_lb_async_$tag python tools/profile_run.py cfg5w 6 0 > gpurun_out/ncu_3.log 2>&1
ncu --set full --clock-control none -k regex:"sk" -s 2 -c 1 -f -o gpurun_out/prof_sk_static_$tag ./tools/microbench/streams2 80485376 ncu1 > gpurun_out/ncu_4.log 2>&1
ncu --set full --clock-control none -k regex:"sk_chunk" -s 2 -c 1 -f -o gpurun_out/prof_sk_dyn_$tag ./tools/microbench/streams2 80485376 ncu2 > gpurun_out/ncu_5.log 2>&1
tail -2 gpurun_out/ncu_*.log
ls -la gpurun_out/*.ncu-rep
