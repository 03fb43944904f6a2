#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for flow in 1; do
echo "== flow=$flow"
CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d_dev.so CHR3D_LN_FLOW=$flow timeout 150 $TR --master-port 29517 scripts/mg_trace.py 2>&1 | grep "^rank" | tee -a gpurun_out/r02_mg8_trace.log
done
