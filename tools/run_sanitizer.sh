mkdir -p gpurun_out
out=gpurun_out/sanitizer_r5p.txt
: > $out
run() { tool=$1; shift; echo "### compute-sanitizer --tool $tool --report-api-errors no :: pytest $*" >> $out; timeout 1500 compute-sanitizer --tool $tool --report-api-errors no --error-exitcode 77 --target-processes all python -m pytest "$@" -m gpu -x -q > $out.raw 2>&1; echo "exit code $?" >> $out; grep -E "ERROR SUMMARY|passed|failed" $out.raw | tail -4 >> $out; }
run memcheck tests/test_exact_kat.py tests/test_multigpu_torchrun.py::test_verify_launch_single_gpu
run memcheck tests/test_gpu_parity.py -k "benchmark_shaped"
run memcheck tests/test_gpu_parity.py -k "strip or asynchronous or in_place or error_codes or device_side"
run memcheck tests/test_multigpu.py
run initcheck tests/test_gpu_parity.py -k "benchmark_shaped and (porous or bcc)"
run racecheck tests/test_gpu_parity.py -k "(lb_steps_bit_exact and (rand33 or slit8)) or (moment_propagation_bit_exact and rand33)"
rm -f $out.raw
cat $out
