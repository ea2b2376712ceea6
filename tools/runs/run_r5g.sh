mkdir -p gpurun_out
{
tools/ab.sh r5g cfg5w 30 "-|" "-|LBG_MP_TPC=1" "-|LBG_MP_TPC=2" "-|LBG_MP_TPC=4" "-|LBG_MP_TPC=8" "-|LBG_MP_TPC=16" "-|LBG_MP_TPC=64" "-|"
tools/ab.sh r5g cfg3 200 "-|" "-|LBG_MP_TPC=1" "-|LBG_MP_TPC=4" "-|LBG_MP_TPC=16"
tools/ab.sh r5g cfg2 400 "-|" "-|LBG_MP_TPC=1" "-|LBG_MP_TPC=4"
} > gpurun_out/ab_r5g.txt 2>&1
cat gpurun_out/ab_r5g.txt
