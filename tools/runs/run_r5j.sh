mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r5j.txt
cat gpurun_out/pytest_r5j.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n2_r5j.json 2> gpurun_out/bench_n2_r5j.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --in-place > gpurun_out/bench_n2_aa_r5j.json 2> gpurun_out/bench_n2_aa_r5j.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --also "" --in-place > gpurun_out/bench_n1_aa_r5j.json 2> gpurun_out/bench_n1_aa_r5j.err
python - <<'PY'
import json
for f in ['bench_n2_r5j','bench_n2_aa_r5j','bench_n1_aa_r5j']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, round(d['value']), d['lb']['ms_per_step'], d['mp']['ms_per_step'], d['config']['phase_a_layout'], d['verify'])
    except Exception as e:
        print(f, 'FAILED', e); print(open('gpurun_out/'+f+'.err').read()[-1500:])
PY
