mkdir -p gpurun_out
LBG_LB_WARP=1 LBG_LB_PIPE=0 LBG_LB_TPC=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lb_steps_bit_exact or benchmark_shaped or lb_strip or medium" 2>&1 | tail -2
{
tools/ab.sh r5v cfg5w 30 "-|" "-|LBG_LB_WARP=1" "-|LBG_LB_WARP=1 LBG_LB_MINB=2" "-|" "-|LBG_LB_WARP=1" "-|LBG_LB_MINB=2"
tools/ab.sh r5v cfg5b 30 "-|" "-|LBG_LB_WARP=1"
tools/ab.sh r5v slitL 30 "-|" "-|LBG_LB_WARP=1"
tools/ab.sh r5v cfg3 200 "-|" "-|LBG_LB_WARP=1 LBG_LB_PIPE=0 LBG_LB_TPC=2 LBG_LB_MINB=3" "-|LBG_LB_WARP=1 LBG_LB_PIPE=0 LBG_LB_TPC=2 LBG_LB_MINB=2"
} > gpurun_out/ab_r5v.txt 2>&1
cat gpurun_out/ab_r5v.txt
