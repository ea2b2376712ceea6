mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py tests/test_multigpu_torchrun.py -m gpu -x -q 2>&1 | tail -3
tools/ncu_one.sh mp_r6c "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r6c "lb_step" cfg5w 8 0
