mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slices or asynchronous" 2>&1 | tail -2
{
tools/ab.sh r5q cfg2 400 "-|" "-|LBG_MP_NBT=1" "-|" "-|LBG_MP_NBT=1"
tools/ab.sh r5q slitL 30 "-|" "-|LBG_MP_NBT=1"
tools/ab.sh r5q cfg3 200 "-|" "-|LBG_MP_NBT=0"
} > gpurun_out/ab_r5q.txt 2>&1
cat gpurun_out/ab_r5q.txt
