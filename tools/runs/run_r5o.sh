mkdir -p gpurun_out
LBG_MP_QPF=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "moment_propagation or benchmark_shaped" 2>&1 | tail -2
tools/ab.sh r5o cfg5w 30 "-|" "-|LBG_MP_QPF=1" "-|" "-|LBG_MP_QPF=1" > gpurun_out/ab_r5o.txt 2>&1
tools/ab.sh r5o cfg3 200 "-|" "-|LBG_MP_QPF=1" >> gpurun_out/ab_r5o.txt 2>&1
cat gpurun_out/ab_r5o.txt
