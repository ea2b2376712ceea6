mkdir -p gpurun_out
{
tools/ab.sh r5s cfg5w 30 "-|" "_ld1|" "_ld2|" "-|" "_ld1|" "_ld2|"
tools/ab.sh r5s cfg3 200 "-|" "_ld1|" "_ld2|"
tools/ab.sh r5s cfg2 400 "-|" "_ld1|" "_ld2|"
} > gpurun_out/ab_r5s.txt 2>&1
cat gpurun_out/ab_r5s.txt
