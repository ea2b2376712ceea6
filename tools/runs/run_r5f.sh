mkdir -p gpurun_out
tools/ncu_one.sh mp_r5f "mp_step_kernel" cfg5w 2 8
tools/ncu_one.sh lb_r5f "lb_step" cfg5w 8 0
ls -la gpurun_out/*.ncu-rep
