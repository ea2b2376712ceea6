mkdir -p gpurun_out
out=gpurun_out/sanitizer_r6f.txt
echo "### compute-sanitizer --tool memcheck --report-api-errors no :: pytest tests -m gpu -k 'not full_size and not torchrun'  (whole single-GPU suite except the full-size lattices)" > $out
( time timeout 2400 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 77 python -m pytest tests -m gpu -x -q -k "not full_size and not process_per_gpu" ) > $out.raw 2>&1
echo "exit code $?" >> $out
grep -E "ERROR SUMMARY|passed|failed|real" $out.raw | tail -5 >> $out
rm -f $out.raw
cat $out
