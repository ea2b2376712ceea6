mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_r5w.txt; cat gpurun_out/pytest_r5w.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_n1_r5w.out 2> gpurun_out/bench_n1_r5w.err ) 2>&1 | grep real
tail -1 gpurun_out/bench_n1_r5w.out > gpurun_out/bench_n1_r5w.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r5w.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['mp_step_kernel']['frac'],3), 'traffic', d['roofline']['traffic'], 'e2e', round(e['value']), {k:round(v,3) for k,v in e['phase_seconds'].items()}, d['verify']['ok'], d['cpu_baseline'], d['clocks'], 'launches', d['gpu_launches'])
print({k:(round(v['value']), round(v['lb_roofline_frac'],3), round(v['mp_roofline_frac'],3)) for k,v in d['also'].items()})
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref_r5w.out 2> gpurun_out/bench_ref_r5w.err ) 2>&1 | grep real
tail -1 gpurun_out/bench_ref_r5w.out | tee gpurun_out/bench_ref_r5w.json | cut -c1-900
