mkdir -p gpurun_out
LBG_MP_QPF=1 python -m pytest tests/test_gpu_parity.py tests/test_exact_kat.py -m gpu -x -q -k "moment_propagation or benchmark_shaped or mp_matches or medium or tuto" 2>&1 | tail -4 > gpurun_out/pytest_r5h.txt
cat gpurun_out/pytest_r5h.txt
{
tools/ab.sh r5h cfg5w 30 "-|" "-|LBG_MP_QPF=1" "-|" "-|LBG_MP_QPF=1"
tools/ab.sh r5h cfg3 200 "-|" "-|LBG_MP_QPF=1"
tools/ab.sh r5h cfg2 400 "-|" "-|LBG_MP_QPF=1"
tools/ab.sh r5h cfg5b 30 "-|" "-|LBG_MP_QPF=1"
} > gpurun_out/ab_r5h.txt 2>&1
cat gpurun_out/ab_r5h.txt
