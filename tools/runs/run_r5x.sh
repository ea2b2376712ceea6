mkdir -p gpurun_out
for i in 1 2; do
LBG_TIMING=1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --also "" > gpurun_out/bench_n1_r5x$i.out 2> gpurun_out/bench_n1_r5x$i.err
tail -1 gpurun_out/bench_n1_r5x$i.out | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v,3) for k,v in d['e2e']['phase_seconds'].items()})"
grep "lbg timing" gpurun_out/bench_n1_r5x$i.err | tail -8
done
