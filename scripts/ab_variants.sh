#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "" _unroll _nohint _unroll_nohint; do
  CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d$v.so timeout 120 python scripts/ab_probe.py 2>&1 | tail -1 | tee -a gpurun_out/r02_ab_variants.log
done
