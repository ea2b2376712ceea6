mkdir -p gpurun_out
{
tools/ab.sh r6g cfg5w 30 "-|" "_st1|" "_st2|" "_ldm2|" "-|" "_st1|" "_st2|" "_ldm2|"
tools/ab.sh r6g cfg3 200 "-|" "_st1|" "_st2|"
tools/ab.sh r6g cfg5b 30 "-|" "_st1|"
} > gpurun_out/ab_r6g.txt 2>&1
cat gpurun_out/ab_r6g.txt
