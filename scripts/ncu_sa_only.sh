#!/bin/bash
# one --set full capture of sa_kernel on the default bench workload (after a plain run of the same command)
TAG=${TAG:-dev}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n > gpurun_out/${TAG}_plain_bench.json 2> gpurun_out/${TAG}_plain_bench.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:sa_kernel -s 3 -c 1 -o gpurun_out/${TAG}_sa_kernel -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n > gpurun_out/${TAG}_ncu_sa.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_sa.log
