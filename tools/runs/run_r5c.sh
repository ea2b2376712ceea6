mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_r5c.txt
cat gpurun_out/pytest_r5c.txt
LBG_TIMING=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --also "" > gpurun_out/bench_n1_r5c.json 2> gpurun_out/bench_n1_r5c.err
tail -c 3000 gpurun_out/bench_n1_r5c.json; grep "lbg timing" gpurun_out/bench_n1_r5c.err | tail -40
LBG_TIMING=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_r5c.json 2> gpurun_out/bench_n2_r5c.err
tail -c 3000 gpurun_out/bench_n2_r5c.json; grep "lbg timing" gpurun_out/bench_n2_r5c.err | tail -60
