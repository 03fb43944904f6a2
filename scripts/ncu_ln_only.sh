#!/bin/bash
# one --set full capture of ln_tile_kernel (after a plain run of the same command)
TAG=${TAG:-dev}
mkdir -p gpurun_out
python scripts/large_n_probe.py 24925 > gpurun_out/${TAG}_plain_large_n.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ln_tile_kernel -s 26 -c 1 -o gpurun_out/${TAG}_ln_tile_kernel -f \
    python scripts/large_n_probe.py 24925 > gpurun_out/${TAG}_ncu_ln.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_ln.log
