#!/bin/bash
# Round profile evidence: launch list of the bench command + one --set full capture of each hot kernel.
# Every ncu run follows a plain run of the same command that exited 0.  Outputs: gpurun_out/r01_*.
set -x
TAG=${TAG:-r01}
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain_bench.json 2> gpurun_out/${TAG}_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sa_kernel -s 3 -c 1 -o gpurun_out/${TAG}_sa_kernel -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n > gpurun_out/${TAG}_ncu_sa.log 2>&1
python scripts/large_n_probe.py 24925 > gpurun_out/${TAG}_plain_large_n.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:ln_tile_kernel -s 3 -c 1 -o gpurun_out/${TAG}_ln_tile_kernel -f \
    python scripts/large_n_probe.py 24925 > gpurun_out/${TAG}_ncu_ln.log 2>&1
ls -la gpurun_out/
