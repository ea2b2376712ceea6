mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py tests/test_multigpu_torchrun.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_r5d.txt
cat gpurun_out/pytest_r5d.txt
LBG_TIMING=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_n1_r5d.json 2> gpurun_out/bench_n1_r5d.err
python -c "
import json;d=json.load(open('gpurun_out/bench_n1_r5d.json'));print('N1',d['value'],d['lb']['ms_per_step'],d['mp']['ms_per_step'],d['e2e']['value'],d['e2e']['phase_seconds'],d['verify']['ok'])"
grep "lbg timing" gpurun_out/bench_n1_r5d.err | tail -8
LBG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_r5d.json 2> gpurun_out/bench_n2_r5d.err
python -c "
import json;d=json.load(open('gpurun_out/bench_n2_r5d.json'));print('N2',d['value'],d['lb']['ms_per_step'],d['mp']['ms_per_step'],d['e2e']['value'],d['e2e']['phase_seconds'],d['verify'])"
grep "lbg timing" gpurun_out/bench_n2_r5d.err | tail -36
