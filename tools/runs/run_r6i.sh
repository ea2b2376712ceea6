mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_exact_kat.py -m gpu -x -q 2>&1 | tail -2
{
tools/ab.sh r6i cfg2 400 "-|LBG_MP_ARITH=0" "-|" "-|LBG_MP_ARITH=0" "-|"
tools/ab.sh r6i slitL 30 "-|LBG_MP_ARITH=0" "-|"
tools/ab.sh r6i cfg3 200 "-|" "-|LBG_MP_NBT=0 LBG_MP_ARITH=0" "-|LBG_MP_NBT=0"
tools/ab.sh r6i cfg4 20 "-|" "-|LBG_MP_NBT=0" 
} > gpurun_out/ab_r6i.txt 2>&1
cat gpurun_out/ab_r6i.txt
