#!/bin/bash
# usage: mkvariant_all.sh name [extra nvcc -D flags...]  (full library, all tiers)
name=$1; shift
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -pthread -I/root/repo/include -I/root/repo/jacobi_pd_b200/csrc "$@" -shared -o /root/repo/tests/scratch/variants/lib_$name.so /root/repo/jacobi_pd_b200/csrc/jpd_api.cu 2>&1 | grep -E "error"
echo built $name
