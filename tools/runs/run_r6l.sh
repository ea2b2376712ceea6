mkdir -p gpurun_out
tools/ncu_one.sh mp_r6l "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r6l "lb_step" cfg5w 8 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5w_r6l.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --also "" --no-verify > gpurun_out/b_r6l.log 2>&1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_r6l.out 2> gpurun_out/bench_n1_r6l.err
tail -1 gpurun_out/bench_n1_r6l.out > gpurun_out/bench_n1_r6l.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r6l.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['mp_step_kernel']['frac'],3), 'e2e', round(e['value']), e['seconds_all'], d['verify']['ok'])
print({k:(round(v['value']), round(v['lb_roofline_frac'],3), round(v['mp_roofline_frac'],3)) for k,v in d['also'].items()})
PY
