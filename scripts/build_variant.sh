#!/bin/bash
# bash scripts/build_variant.sh <name> <extra nvcc flags...>: chromosome3d_b200/libchr3d_<name>.so with large_n.cu recompiled under the flags
# (development: A/B runs of kernel variants in one GPU visit, selected with CHR3D_LIBRARY)
set -e
cd "$(dirname "$0")/../chromosome3d_b200/csrc"
name=$1; shift
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr -DCHR3D_WITH_NCCL"
$NV "$@" -c -o large_n_$name.o large_n.cu 2> large_n_$name.ptxas.log || { cat large_n_$name.ptxas.log; exit 1; }
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libchr3d_$name.so api.o restraints.o sa_ensemble.o assess.o microbench.o large_n_$name.o tools.o dg.o -Xcompiler -fPIC -ldl
echo built libchr3d_$name.so
