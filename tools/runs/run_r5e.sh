mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_r5e.txt
cat gpurun_out/pytest_r5e.txt
{
tools/ab.sh r5e cfg5w 30 "-|" "-|LBG_BENCH_KA=0" "-|"
tools/ab.sh r5e cfg3 200 "-|"
tools/ab.sh r5e cfg2 400 "-|"
tools/ab.sh r5e cfg5b 30 "-|"
} > gpurun_out/ab_r5e.txt 2>&1
cat gpurun_out/ab_r5e.txt
