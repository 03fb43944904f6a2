#!/bin/bash
# ncu --set full capture of the SA kernel (shortened schedule; never a bench value)
export CHR3D_SA_CLUSTER=${C:-6} CHR3D_SA_WARPS=${W:-12}
python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:sa_kernel -c 1 -o gpurun_out/prof_sa2 -f python bench.py --profile --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_sa2.log 2>&1
tail -3 gpurun_out/ncu_sa2.log
