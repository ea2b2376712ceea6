mkdir -p gpurun_out
{
tools/ab.sh r5u cfg5w 30 "-|" "-|LBG_MP_SYNC=1" "-|" "-|LBG_MP_SYNC=1"
tools/ab.sh r5u cfg3 200 "-|" "-|LBG_MP_SYNC=1"
tools/ab.sh r5u cfg2 400 "-|" "-|LBG_MP_SYNC=1"
tools/ab.sh r5u slitL 30 "-|" "-|LBG_MP_SYNC=1"
} > gpurun_out/ab_r5u.txt 2>&1
cat gpurun_out/ab_r5u.txt
