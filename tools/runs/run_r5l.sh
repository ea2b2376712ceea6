mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_r5l.txt
cat gpurun_out/pytest_r5l.txt
LBG_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1_r5l.out 2> gpurun_out/bench_n1_r5l.err
tail -1 gpurun_out/bench_n1_r5l.out > gpurun_out/bench_n1_r5l.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_r5l.json').read())
e=d['e2e']
print('N1', round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['mp_step_kernel']['frac'], 'e2e', round(e['value']), {k:round(v,3) for k,v in e['phase_seconds'].items()}, d['verify']['ok'], d.get('cpu_baseline'), d.get('also'))
PY
grep "lbg timing" gpurun_out/bench_n1_r5l.err | tail -9
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_r5l.out 2> gpurun_out/bench_ref_r5l.err; tail -1 gpurun_out/bench_ref_r5l.out; free -g | head -2; nproc
