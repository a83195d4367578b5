#!/bin/bash
# Kernel-variant experiments: build variants/libepid_ncw<N>.so with N consumer warps in k_pf_stream (select with EPID_LIB=...).
set -e
cd "$(dirname "$0")/../pylinac_b200/csrc"
make -j8 >/dev/null
mkdir -p ../../variants
for n in "$@"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -cudart static \
       -DEPID_ST_NCW=$n -c pf_stream.cu -o /tmp/pf_stream_$n.o
  objs=$(ls build/*.o | grep -v pf_stream.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../../variants/libepid_ncw$n.so $objs /tmp/pf_stream_$n.o -ldl
done
