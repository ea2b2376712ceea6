mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r5m.txt
cat gpurun_out/pytest_r5m.txt
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_n1_r5m.out 2> gpurun_out/bench_n1_r5m.err
tail -1 gpurun_out/bench_n1_r5m.out > gpurun_out/bench_n1_r5m.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r5m.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['mp_step_kernel']['frac'], 'e2e', round(e['value']), {k:round(v,3) for k,v in e['phase_seconds'].items()}, d['verify']['ok'])
PY
tools/ncu_one.sh mp_r5m "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r5m "lb_step" cfg5w 8 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5w_r5m.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --also "" --no-verify > gpurun_out/b_r5m.log 2>&1
