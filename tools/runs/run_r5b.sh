mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r5b.txt
{
tools/ab.sh r5b cfg5w 30 "_r1|" "-|" "-|LBG_LB_TPC=1" "-|LBG_LB_TPC=4" "_r1|"
tools/ab.sh r5b cfg3 200 "_r1|" "-|"
tools/ab.sh r5b cfg2 400 "_r1|" "-|"
} > gpurun_out/ab_r5b.txt 2>&1
cat gpurun_out/pytest_r5b.txt gpurun_out/ab_r5b.txt
