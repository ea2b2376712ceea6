mkdir -p gpurun_out
LBG_LB_STRIP_ROWS=3 LBG_LB_PIPE=0 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lb_steps_bit_exact or benchmark_shaped or medium or equilibration" 2>&1 | tail -4 > gpurun_out/pytest_r5i.txt
LBG_LB_STRIP_ROWS=2 LBG_LB_PIPE=0 LBG_LB_TPC=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lb_steps_bit_exact or benchmark_shaped" 2>&1 | tail -4 >> gpurun_out/pytest_r5i.txt
cat gpurun_out/pytest_r5i.txt
{
tools/ab.sh r5i cfg5w 30 "-|" "-|LBG_LB_STRIP_ROWS=32" "-|LBG_LB_STRIP_ROWS=64" "-|LBG_LB_STRIP_ROWS=128" "-|LBG_LB_STRIP_ROWS=256" "-|LBG_LB_STRIP_ROWS=512" "-|"
tools/ab.sh r5i cfg5b 30 "-|" "-|LBG_LB_STRIP_ROWS=64" "-|LBG_LB_STRIP_ROWS=128"
tools/ab.sh r5i slitL 30 "-|" "-|LBG_LB_STRIP_ROWS=128"
} > gpurun_out/ab_r5i.txt 2>&1
cat gpurun_out/ab_r5i.txt
