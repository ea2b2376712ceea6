mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_r5a.txt
{
tools/ab.sh r5a cfg5w 30 "_r1|" "-|LBG_LB_PIPE=0" "-|LBG_LB_PIPE=1" "-|LBG_LB_PIPE=0 LBG_LB_TPC=2 LBG_LB_MINB=3" "-|LBG_LB_PIPE=0 LBG_LB_TPC=2" "-|LBG_LB_PIPE=0 LBG_LB_MINB=3" "-|LBG_LB_PIPE=0 LBG_MP_TPC=16" "_r1|"
tools/ab.sh r5a cfg3 200 "_r1|" "-|LBG_LB_PIPE=0" "-|LBG_LB_PIPE=1 LBG_LB_MINB=2" "-|LBG_LB_PIPE=1 LBG_LB_MINB=3"
tools/ab.sh r5a cfg2 400 "_r1|" "-|LBG_LB_PIPE=0" "-|LBG_LB_PIPE=1 LBG_LB_MINB=2" "-|LBG_LB_PIPE=1 LBG_LB_MINB=3"
} > gpurun_out/ab_r5a.txt 2>&1
cat gpurun_out/pytest_r5a.txt gpurun_out/ab_r5a.txt
