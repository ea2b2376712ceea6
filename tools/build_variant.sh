#!/bin/bash
# usage: tools/build_variant.sh <suffix> <nvcc -D flags...>   -- builds laboetie_b200/lib/liblaboetie_gpu<suffix>.so for A/B tuning runs
set -e
suf=$1; shift
cd "$(dirname "$0")/../laboetie_b200/csrc"
mkdir -p /tmp/lbgv$suf
for f in api geometry lb_kernels lb_aa_kernels mp_kernels; do
  /usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ccbin /usr/bin/g++ "$@" -c $f.cu -o /tmp/lbgv$suf/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/liblaboetie_gpu$suf.so /tmp/lbgv$suf/*.o -cudart static -ldl -lpthread -lrt
ls -la ../lib/liblaboetie_gpu$suf.so
