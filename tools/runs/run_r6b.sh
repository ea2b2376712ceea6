mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 > gpurun_out/pytest_r6b.txt; cat gpurun_out/pytest_r6b.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_r6b.out 2> gpurun_out/bench_n1_r6b.err
tail -1 gpurun_out/bench_n1_r6b.out > gpurun_out/bench_n1_r6b.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r6b.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['mp_step_kernel']['frac'],3), 'traffic', d['roofline']['traffic'], d['roofline']['mp_step_kernel'].get('traffic'), 'e2e', round(e['value']), e['seconds_all'], d['verify']['ok'], d['cpu_baseline']['value'], list(d.keys()))
PY
