mkdir -p gpurun_out
{
tools/ab.sh r5r cfg5w 30 "-|" "_mp1|" "_mp3|" "-|"
tools/ab.sh r5r cfg3 200 "-|" "_mp1|" "_mp3|"
} > gpurun_out/ab_r5r.txt 2>&1
cat gpurun_out/ab_r5r.txt
( time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size_config5" ) 2>&1 | tail -8
