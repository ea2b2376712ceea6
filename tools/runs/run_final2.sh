mkdir -p gpurun_out
tools/ncu_one.sh mp_r5y "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r5y "lb_step" cfg5w 8 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5w_r5y.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --also "" --no-verify > gpurun_out/b_r5y.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_r5y.out 2> gpurun_out/bench_n1_r5y.err
tail -1 gpurun_out/bench_n1_r5y.out > gpurun_out/bench_n1_r5y.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r5y.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['mp_step_kernel']['frac'],3), 'e2e', round(e['value']), e['seconds_all'], {k:round(v,3) for k,v in e['phase_seconds'].items()}, d['verify']['ok'])
PY
