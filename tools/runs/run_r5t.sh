mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_exact_kat.py -m gpu -x -q 2>&1 | tail -2
{
tools/ab.sh r5t cfg5w 30 "_prev|" "-|" "_prev|" "-|" "-|LBG_LB_TPC=1" "-|LBG_LB_TPC=4"
tools/ab.sh r5t cfg5w 30 "_prev|--in-place" 
} > gpurun_out/ab_r5t.txt 2>&1
LBG_LIB=$PWD/laboetie_b200/lib/liblaboetie_gpu_prev.so python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --also "" --in-place 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev in-place lb', d['lb']['ms_per_step'])" >> gpurun_out/ab_r5t.txt
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --also "" --in-place 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new in-place lb', d['lb']['ms_per_step'])" >> gpurun_out/ab_r5t.txt
cat gpurun_out/ab_r5t.txt
