mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5 > gpurun_out/pytest_r6j.txt; cat gpurun_out/pytest_r6j.txt
tools/ncu_one.sh mp_r6j "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r6j "lb_step" cfg5w 8 0
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_r6j.out 2> gpurun_out/bench_n1_r6j.err
tail -1 gpurun_out/bench_n1_r6j.out > gpurun_out/bench_n1_r6j.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r6j.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['mp_step_kernel']['frac'],3), 'traffic', d['roofline']['traffic'], 'e2e', round(e['value']), e['seconds_all'], d['verify']['ok'])
print({k:(round(v['value']), round(v['lb_roofline_frac'],3), round(v['mp_roofline_frac'],3)) for k,v in d['also'].items()})
PY
