mkdir -p gpurun_out
for i in 1 2 3; do
LBG_TIMING=1 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --also "" --no-verify > gpurun_out/bench_n1_r6k$i.out 2> gpurun_out/bench_n1_r6k$i.err
tail -1 gpurun_out/bench_n1_r6k$i.out | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['seconds_all'], {k:round(v,3) for k,v in d['e2e']['phase_seconds'].items()})"
grep "lbg_mp_init" gpurun_out/bench_n1_r6k$i.err | tail -6
done
