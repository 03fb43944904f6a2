#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 24925 4985; do
echo "== N=$n"
TRACE_N=$n RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29519 CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d_dev.so CHR3D_LN_FLOW=1 timeout 150 python scripts/mg_trace.py 2>&1 | grep "^rank" | tee -a gpurun_out/r02_trace1.log
done
